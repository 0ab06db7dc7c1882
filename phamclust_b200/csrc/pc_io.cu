// pc_io.cu -- host-side text formats of the matrix (no GPU work): the
// squareform / lower-triangle / adjacency writers and the squareform reader of
// /root/reference/src/phamclust/matrix.py:500-670, operating directly on the
// packed int32 micro-unit triangle ("%.6f" of n/1e6 is pure integer -> text).
// without this the hand-off of a 10k-genome matrix spends
// minutes in per-value Python formatting.  The lower-triangle form is the
// pipeline's on-disk cache (scripts/phamclust.py:254-268) and stays byte-identical.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/phamclust_b200.h"

static thread_local std::string g_io_error;

extern "C" const char *pc_io_last_error(void) { return g_io_error.c_str(); }

static inline int64_t tri(int64_t n, int64_t i, int64_t j)
{
    if (i > j) { const int64_t t = i; i = j; j = t; }
    return i * n - (i * (i - 1)) / 2 + (j - i);
}

// "%.6f" of micro/1e6 for 0 <= micro <= 1e6 (the only values a SymMatrix holds)
static inline char *put_micro(char *p, int32_t micro)
{
    *p++ = (char)('0' + micro / 1000000);
    *p++ = '.';
    int32_t frac = micro % 1000000;
    for (int d = 100000; d >= 1; d /= 10) { *p++ = (char)('0' + frac / d); frac %= d; }
    return p;
}

// mode 0: squareform with header names, 1: lower triangle, 2: adjacency, 3: adjacency skipping zeros.
// order[k] = storage index of the k-th node in traversal order; names are a blob + offsets (storage order).
extern "C" int pc_write_matrix(const char *path, int64_t n, const int32_t *micro, const int64_t *order,
                               const char *names, const int64_t *name_off, int mode)
{
    if (!path || n < 0 || (n > 0 && (!micro || !order || !names || !name_off)) || mode < 0 || mode > 3) {
        g_io_error = "pc_write_matrix: bad arguments";
        return PC_EINVAL;
    }
    for (int64_t k = 0; k < n; ++k)
        if (order[k] < 0 || order[k] >= n) { g_io_error = "pc_write_matrix: bad traversal order"; return PC_EINVAL; }
    FILE *f = fopen(path, "wb");
    if (!f) { g_io_error = std::string("cannot open ") + path; return PC_EINVAL; }
    std::vector<char> buf;
    auto name_of = [&](int64_t k, const char *&s, size_t &len) {
        const int64_t g = order[k];
        s = names + name_off[g];
        len = (size_t)(name_off[g + 1] - name_off[g]);
    };
    if (mode <= 1) {
        std::string header = std::to_string(n);
        if (mode == 0)
            for (int64_t k = 0; k < n; ++k) {
                const char *s; size_t len;
                name_of(k, s, len);
                header += '\t';
                header.append(s, len);
            }
        header += '\n';
        fwrite(header.data(), 1, header.size(), f);
    }
    for (int64_t i = 0; i < n; ++i) {
        const char *s; size_t len;
        name_of(i, s, len);
        if (mode <= 1) {
            const int64_t cols = mode == 1 ? i + 1 : n;
            buf.resize(len + (size_t)cols * 9 + 2);
            char *p = buf.data();
            memcpy(p, s, len); p += len;
            for (int64_t j = 0; j < cols; ++j) {
                const int32_t v = micro[tri(n, order[i], order[j])];
                if (v < 0 || v > 1000000) { fclose(f); g_io_error = "unset or out-of-range weight"; return PC_EINVAL; }
                *p++ = '\t';
                p = put_micro(p, v);
            }
            *p++ = '\n';
            fwrite(buf.data(), 1, (size_t)(p - buf.data()), f);
        } else {
            for (int64_t j = i; j < n; ++j) {
                const int32_t v = micro[tri(n, order[i], order[j])];
                if (v < 0 || v > 1000000) { fclose(f); g_io_error = "unset or out-of-range weight"; return PC_EINVAL; }
                if (mode == 3 && v == 0) continue;
                const char *t; size_t tl;
                name_of(j, t, tl);
                buf.resize(len + tl + 12);
                char *p = buf.data();
                memcpy(p, s, len); p += len; *p++ = '\t';
                memcpy(p, t, tl); p += tl; *p++ = '\t';
                p = put_micro(p, v);
                *p++ = '\n';
                fwrite(buf.data(), 1, (size_t)(p - buf.data()), f);
            }
        }
    }
    if (fclose(f) != 0) { g_io_error = "write failed"; return PC_EINVAL; }
    return PC_OK;
}

// CPython's round(x, 6) * 1e6 on the host (same recipe as the device epilogue)
static inline long long round6_micro_host(double x)
{
    const double hi = x * 1e6;
    const double lo = std::fma(x, 1e6, -hi);
    double r = std::nearbyint(hi);
    if (std::fabs(hi - r) == 0.5) {
        const double fl = std::floor(hi);
        if (lo > 0.0) r = fl + 1.0; else if (lo < 0.0) r = fl;
    }
    return (long long)r;
}

// Parse a squareform / lower-triangle file held in memory (matrix.py:564-621):
// line 0 = size (+ names, ignored); row i = name TAB v0 TAB v1 ...; only the
// first i+1 values of row i are used.  micro_out is the packed triangle in FILE
// order; name_begin/name_end give each row name's byte range in buf.
extern "C" int pc_parse_squareform(const char *buf, int64_t len, int64_t n, int32_t *micro_out,
                                   int64_t *name_begin, int64_t *name_end)
{
    if (!buf || n < 0 || (n > 0 && (!micro_out || !name_begin || !name_end))) {
        g_io_error = "pc_parse_squareform: bad arguments";
        return PC_EINVAL;
    }
    int64_t pos = 0;
    while (pos < len && buf[pos] != '\n') ++pos;      // skip the size line
    ++pos;
    for (int64_t i = 0; i < n; ++i) {
        if (pos >= len) { g_io_error = "squareform file has fewer rows than its header says"; return PC_EINVAL; }
        int64_t eol = pos;
        while (eol < len && buf[eol] != '\n') ++eol;
        int64_t p = pos;
        while (p < eol && buf[p] != '\t') ++p;
        name_begin[i] = pos;
        name_end[i] = p;
        for (int64_t j = 0; j <= i; ++j) {
            if (p >= eol) { g_io_error = "squareform row is too short"; return PC_EINVAL; }
            ++p;                                      // the tab
            // the field is buf[p, q): up to the next tab or the end of THIS line; it is copied out
            // so that strtod can neither run into the next row (it skips leading whitespace,
            // newlines included) nor accept what float() refuses for a weight (nan, inf, hex)
            int64_t q = p;
            while (q < eol && buf[q] != '\t') ++q;
            char field[64];
            const int64_t flen = q - p;
            bool plain = flen > 0 && flen < (int64_t)sizeof field;
            for (int64_t k = 0; plain && k < flen; ++k) {
                const char c = buf[p + k];
                plain = (c >= '0' && c <= '9') || c == '.' || c == '-' || c == '+' || c == 'e' || c == 'E' ||
                        (c == '\r' && k == flen - 1);
            }
            if (!plain) { g_io_error = "could not convert string to float"; return PC_EINVAL; }
            memcpy(field, buf + p, (size_t)flen);
            field[flen] = 0;
            if (field[flen - 1] == '\r') field[flen - 1] = 0;
            char *endp = nullptr;
            const double v = strtod(field, &endp);     // digits, sign, '.', exponent only: locale-proof for '.'
            if (endp == field || *endp != 0) { g_io_error = "could not convert string to float"; return PC_EINVAL; }
            p = q;
            if (!(v >= 0.0 && v <= 1.0)) { g_io_error = "weight not in [0.0, 1.0]"; return PC_EINVAL; }
            micro_out[tri(n, j, i)] = (int32_t)round6_micro_host(v);
        }
        pos = eol + 1;
    }
    return PC_OK;
}

// One pass over a 2/3-column TSV held in memory (scripts/phamclust.py:21-47):
// per row the byte ranges of name / pham / translation (translation range empty
// for 2-column rows -> "M").  Returns the number of rows, or -1 on a malformed
// row (anything but 2 or 3 columns after rstrip()).
extern "C" int64_t pc_scan_tsv(const char *buf, int64_t len, int64_t max_rows, int64_t *fields /* 6 per row */)
{
    int64_t rows = 0, pos = 0;
    while (pos < len) {
        int64_t eol = pos;
        while (eol < len && buf[eol] != '\n') ++eol;
        int64_t end = eol;                            // rstrip(): trailing whitespace incl. tabs
        while (end > pos && (buf[end - 1] == ' ' || buf[end - 1] == '\t' || buf[end - 1] == '\r' ||
                             buf[end - 1] == '\n' || buf[end - 1] == '\v' || buf[end - 1] == '\f'))
            --end;
        int64_t t1 = -1, t2 = -1, tabs = 0;
        for (int64_t p = pos; p < end; ++p)
            if (buf[p] == '\t') { if (tabs == 0) t1 = p; else if (tabs == 1) t2 = p; ++tabs; }
        if (tabs < 1 || tabs > 2) { g_io_error = "input file must either 2 or 3 columns"; return -1; }
        if (rows < max_rows) {
            int64_t *f = fields + 6 * rows;
            f[0] = pos; f[1] = t1;
            f[2] = t1 + 1; f[3] = tabs == 2 ? t2 : end;
            f[4] = tabs == 2 ? t2 + 1 : end; f[5] = end;
            if (tabs == 1) f[4] = f[5] = -1;
        }
        ++rows;
        pos = eol + 1;
    }
    return rows;
}

// Intern one column of the scanned TSV: ids[r] = index of the distinct byte string
// buf[fields[6r + 2*col] .. fields[6r + 2*col + 1]) in FIRST-SEEN order (the order in which
// load_genomes_from_tsv's dict meets genome names, scripts/phamclust.py:41-45);
// first_row[id] = the row that introduced it.  col 0 = name, 1 = pham.  Returns the number
// of distinct strings (first_row must hold `rows` entries at most), or -1.
extern "C" int64_t pc_tsv_intern(const char *buf, int64_t rows, const int64_t *fields, int col,
                                 int32_t *ids, int64_t *first_row)
{
    if (rows < 0 || (rows > 0 && (!buf || !fields || !ids || !first_row)) || col < 0 || col > 1) {
        g_io_error = "pc_tsv_intern: bad arguments";
        return -1;
    }
    std::unordered_map<std::string_view, int32_t> seen;
    seen.reserve((size_t)std::min<int64_t>(rows, 1 << 20));
    for (int64_t r = 0; r < rows; ++r) {
        const int64_t b = fields[6 * r + 2 * col], e = fields[6 * r + 2 * col + 1];
        const std::string_view key(buf + b, (size_t)(e - b));
        auto it = seen.find(key);
        if (it == seen.end()) {
            const int32_t id = (int32_t)seen.size();
            first_row[id] = r;
            it = seen.emplace(key, id).first;
        }
        ids[r] = it->second;
    }
    return (int64_t)seen.size();
}

// Concatenate n byte ranges of buf into out (the residue arena): range k is
// buf[start[k] .. start[k] + len[k]), or the single byte `fill` when start[k] < 0
// (2-column rows carry the translation "M", scripts/phamclust.py:37-38).
extern "C" int pc_gather_fields(const char *buf, int64_t n, const int64_t *start, const int64_t *len,
                                uint8_t fill, uint8_t *out)
{
    if (n < 0 || (n > 0 && (!buf || !start || !len || !out))) {
        g_io_error = "pc_gather_fields: bad arguments";
        return PC_EINVAL;
    }
    int64_t at = 0;
    for (int64_t k = 0; k < n; ++k) {
        if (start[k] < 0) { out[at++] = fill; continue; }
        memcpy(out + at, buf + start[k], (size_t)len[k]);
        at += len[k];
    }
    return PC_OK;
}

// Dense square view of the packed triangle in traversal order (SymMatrix.to_ndarray,
// /root/reference/src/phamclust/matrix.py:155-167: what clustering.py:27 feeds to scikit-learn).
// kind 0: float64 weights micro / 1e6 (the exact value round(w, 6) stores); every entry must
// be a set weight in [0, 1e6] (PC_EINVAL otherwise: the caller falls back to the object
// array the reference builds for unset entries).  kind 1: the int32 micro-units as they are.
// order[k] = storage index of the k-th node in traversal order.
extern "C" int pc_dense_from_micro(int64_t n, const int32_t *micro, const int64_t *order, int kind,
                                   void *out)
{
    if (n < 0 || (n > 0 && (!micro || !order || !out)) || kind < 0 || kind > 1) {
        g_io_error = "pc_dense_from_micro: bad arguments";
        return PC_EINVAL;
    }
    for (int64_t k = 0; k < n; ++k)
        if (order[k] < 0 || order[k] >= n) { g_io_error = "pc_dense_from_micro: bad traversal order"; return PC_EINVAL; }
    // idx(a, b >= a) = base[a] + b
    std::vector<int64_t> base((size_t)n);
    for (int64_t a = 0; a < n; ++a) base[a] = a * n - (a * (a - 1)) / 2 - a;
    const unsigned hw = std::thread::hardware_concurrency();
    const int T = n < 512 ? 1 : (int)std::max(1u, std::min(8u, hw));
    std::vector<int> bad((size_t)T, 0);
    // 64 x 64 tiles; inside a tile the loop whose index runs along the packed rows of the
    // triangle goes innermost (columns above the diagonal, rows below it), so both halves
    // read the triangle contiguously when the traversal order is (piecewise) the storage order
    constexpr int64_t TILE = 64;
    const int64_t n_blocks = (n + TILE - 1) / TILE;
    auto body = [&](int t) {
        double *outd = static_cast<double *>(out);
        int32_t *outi = static_cast<int32_t *>(out);
        auto put = [&](int64_t k, int64_t l) {
            const int64_t a = order[k], b = order[l];
            const int32_t v = micro[b >= a ? base[a] + b : base[b] + a];
            if (kind == 0) {
                if ((uint32_t)v > 1000000u) bad[t] = 1;
                outd[k * n + l] = (double)v / 1e6;
            } else {
                outi[k * n + l] = v;
            }
        };
        for (int64_t kb = n_blocks * t / T; kb < n_blocks * (t + 1) / T; ++kb) {
            const int64_t k0 = kb * TILE, k1 = std::min(n, k0 + TILE);
            for (int64_t l0 = 0; l0 < n; l0 += TILE) {
                const int64_t l1 = std::min(n, l0 + TILE);
                if (order[l0] >= order[k0]) {
                    for (int64_t k = k0; k < k1; ++k)
                        for (int64_t l = l0; l < l1; ++l) put(k, l);
                } else {
                    for (int64_t l = l0; l < l1; ++l)
                        for (int64_t k = k0; k < k1; ++k) put(k, l);
                }
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < T; ++t) pool.emplace_back(body, t);
    body(0);
    for (auto &th : pool) th.join();
    for (int t = 0; t < T; ++t)
        if (bad[t]) { g_io_error = "pc_dense_from_micro: unset or out-of-range weight"; return PC_EINVAL; }
    return PC_OK;
}


// Per node, the sum of every weight it appears in while iterating the upper half of the
// matrix -- the diagonal counts twice, exactly the lists SymMatrix.medoid builds upstream
// (matrix.py:81-96) -- as exact integers in micro-units, in STORAGE order, in one pass over the
// packed triangle (no dense N x N).  PC_EINVAL if a weight is unset / out of range.
extern "C" int pc_node_sums(int64_t n, const int32_t *micro, int64_t *sums)
{
    if (n < 0 || (n > 0 && (!micro || !sums))) {
        g_io_error = "pc_node_sums: bad arguments";
        return PC_EINVAL;
    }
    const unsigned hw = std::thread::hardware_concurrency();
    const int T = n < 2048 ? 1 : (int)std::max(1u, std::min(8u, hw));
    std::vector<std::vector<int64_t>> part((size_t)T);
    std::vector<int> bad((size_t)T, 0);
    // rows are dealt so that every thread sees about the same number of entries
    auto row_of = [&](int t) {
        const double frac = (double)t / T;
        return (int64_t)std::llround((double)n * (1.0 - std::sqrt(1.0 - frac)));
    };
    auto body = [&](int t) {
        std::vector<int64_t> &acc = part[t];
        acc.assign((size_t)n, 0);
        const int64_t r0 = t == 0 ? 0 : row_of(t), r1 = t == T - 1 ? n : row_of(t + 1);
        for (int64_t i = r0; i < r1; ++i) {
            const int32_t *row = micro + (i * n - (i * (i - 1)) / 2);
            int64_t own = 0;
            for (int64_t j = i; j < n; ++j) {
                const int32_t v = row[j - i];
                if ((uint32_t)v > 1000000u) bad[t] = 1;
                own += v;
                acc[j] += v;                 // as target (j == i: the diagonal's second appearance)
            }
            acc[i] += own;                   // as source
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < T; ++t) pool.emplace_back(body, t);
    body(0);
    for (auto &th : pool) th.join();
    for (int t = 0; t < T; ++t)
        if (bad[t]) { g_io_error = "pc_node_sums: unset or out-of-range weight"; return PC_EINVAL; }
    for (int64_t k = 0; k < n; ++k) {
        int64_t v = 0;
        for (int t = 0; t < T; ++t) v += part[t][k];
        sums[k] = v;
    }
    return PC_OK;
}

// One single-genome FASTA file held in memory (fasta.py:4-40 + genome.py:66-83): per record
// the pham named by the first `pham=` field of the header (fields split on '|', each exactly one
// '=') and the sequence with every line rstrip()ped and concatenated.  Pham names are appended to
// name_out, sequences to seq_out; fields[4r..] = name begin/end (in name_out), seq begin/end (in
// seq_out).  Lines end at '\n', '\r\n' or a lone '\r' (Python's universal newlines).  Returns the
// number of records or: -1 the file does not start with '>', -2 a header field without exactly one
// '=', -3 no pham field, -4 out of capacity, -5 a sequence byte outside ASCII (the text-mode
// reader upstream would decode it to a character that does not fit one arena byte).
extern "C" int64_t pc_scan_fasta(const char *buf, int64_t len, int64_t max_records, int64_t *fields,
                                 uint8_t *name_out, int64_t name_cap, uint8_t *seq_out, int64_t seq_cap,
                                 int64_t *used /* [2]: name bytes, seq bytes */)
{
    auto is_space = [](char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; };
    int64_t pos = 0, rec = -1, n_name = 0, n_seq = 0;
    if (len == 0) { used[0] = used[1] = 0; return 0; }    // upstream: next() on an empty file raises; the caller decides
    while (pos < len) {
        int64_t eol = pos;
        while (eol < len && buf[eol] != '\n' && buf[eol] != '\r') ++eol;
        int64_t next = eol + 1;
        if (eol < len && buf[eol] == '\r' && eol + 1 < len && buf[eol + 1] == '\n') next = eol + 2;
        int64_t end = eol;
        while (end > pos && is_space(buf[end - 1])) --end;
        if (buf[pos] == '>') {
            if (++rec >= max_records) { g_io_error = "pc_scan_fasta: too many records"; return -4; }
            // header = line[1:].rstrip(); fields split on '|'
            int64_t f0 = pos + 1;
            bool found = false;
            const int64_t hend = std::max(end, f0);
            for (int64_t p = f0; !found;) {
                int64_t q = p;
                while (q < hend && buf[q] != '|') ++q;
                int64_t eq = -1, n_eq = 0;
                for (int64_t k = p; k < q; ++k)
                    if (buf[k] == '=') { if (eq < 0) eq = k; ++n_eq; }
                if (n_eq != 1) { g_io_error = "FASTA header fields must look like key=value"; return -2; }
                if (eq - p == 4 && memcmp(buf + p, "pham", 4) == 0) {
                    const int64_t vlen = q - (eq + 1);
                    if (n_name + vlen > name_cap) { g_io_error = "pc_scan_fasta: name buffer too small"; return -4; }
                    memcpy(name_out + n_name, buf + eq + 1, (size_t)vlen);
                    fields[4 * rec + 0] = n_name;
                    fields[4 * rec + 1] = n_name + vlen;
                    n_name += vlen;
                    found = true;
                }
                if (q >= hend) break;
                p = q + 1;
            }
            if (!found) { g_io_error = "unable to get pham from FASTA header"; return -3; }
            fields[4 * rec + 2] = fields[4 * rec + 3] = n_seq;
        } else {
            if (rec < 0) { g_io_error = "records in FASTA files must start with '>'"; return -1; }
            const int64_t slen = end - pos;
            if (n_seq + slen > seq_cap) { g_io_error = "pc_scan_fasta: sequence buffer too small"; return -4; }
            for (int64_t k = 0; k < slen; ++k) {
                const unsigned char c = (unsigned char)buf[pos + k];
                if (c >= 0x80) { g_io_error = "translations must be made of single-byte characters"; return -5; }
                seq_out[n_seq + k] = c;
            }
            n_seq += slen;
            fields[4 * rec + 3] = n_seq;
        }
        pos = next;
    }
    used[0] = n_name;
    used[1] = n_seq;
    return rec + 1;
}
