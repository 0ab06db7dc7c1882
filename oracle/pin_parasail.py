"""oracle/pin_parasail.py -- TEST INFRASTRUCTURE: pin the restated tie rules against a
REAL parasail, the moment one is importable.

The oracle restates ``parasail.nw_trace_diag_16(a, b, 11, 1, blosum62).get_traceback("|", "+",
" ")`` (/root/reference/src/phamclust/metrics.py:160-175) from its published algorithm; parasail
itself (PyPI ``parasail~=1.3.0``, requirements.txt:5) is not installable in the build container, so
the rules U2 (H tie priority), U3 (open-vs-extend ties) and U5 (case) of SURVEY.md Appendix A are
UNPINNED.  This script closes that gap wherever the wheel exists:

    python -m oracle.pin_parasail            # exit 0: committed rules == parasail on every case
                                             # exit 1: they differ (the matching combination is printed)
                                             # exit 2: no real parasail importable

It runs the golden alignment vectors plus tie-rich generated pairs through the real library and
through the pure-Python twin (oracle/nw_py.py) under every combination of the rule switches, and
compares the three traceback strings.  ``tests/test_parasail_pin.py`` runs the same comparison and
skips when the library is absent.
"""
import importlib
import itertools
import json
import pathlib
import random
import sys

_HERE = pathlib.Path(__file__).resolve().parent
_STUB = _HERE / "parasail_stub"


def real_parasail():
    """The installed parasail module, or None (never the stub in oracle/parasail_stub)."""
    saved_path = list(sys.path)
    saved_mod = sys.modules.pop("parasail", None)
    try:
        sys.path[:] = [p for p in sys.path if pathlib.Path(p or ".").resolve() != _STUB]
        try:
            mod = importlib.import_module("parasail")
        except Exception:
            return None
        origin = pathlib.Path(getattr(mod, "__file__", "") or "").resolve()
        if _STUB in origin.parents or not hasattr(mod, "nw_trace_diag_16") or not hasattr(mod, "blosum62"):
            return None
        return mod
    finally:
        sys.path[:] = saved_path
        sys.modules.pop("parasail", None)
        if saved_mod is not None:
            sys.modules["parasail"] = saved_mod


def cases(n_random=600, seed=2718, max_len=10 ** 9):
    """(query, ref) pairs: the committed golden vectors + tie-rich random pairs (two- and
    three-letter alphabets make co-optimal alignments the rule), mixed case and symbols
    outside the matrix alphabet for U5/U6."""
    from oracle import testdata
    out = []
    golden = _HERE.parent / "tests" / "golden" / "alignments.json"
    if golden.exists():
        out += [(c["query"], c["ref"]) for c in json.loads(golden.read_text())
                if c["query"] and c["ref"] and max(len(c["query"]), len(c["ref"])) <= max_len]
    rng = random.Random(seed)
    for t in range(n_random):
        alpha = ["AG", "AGS", "NDB", "WC", testdata.AA][t % 5]
        a = testdata.random_protein(rng, 1, 60, alpha)
        b = testdata.mutate(rng, a, rng.random() * 0.7, alpha) if t % 2 else testdata.random_protein(rng, 1, 60, alpha)
        if b:
            out.append((a, b))
    out += [("mktwACDE", "MKTWacde"), ("MKUXO*BZ", "MKOXU*ZB"), ("JJJJ", "jJjJ")]
    return out


def compare(parasail, pairs=None):
    """{(U2 diag>del>ins, U3 ties extend, U5 case sensitive): mismatching case count}."""
    from oracle import nw_py
    pairs = pairs if pairs is not None else cases()
    want = []
    for a, b in pairs:
        tb = parasail.nw_trace_diag_16(a, b, 11, 1, parasail.blosum62).get_traceback("|", "+", " ")
        want.append((tb.query, tb.comp, tb.ref))
    saved = (nw_py.H_PRIORITY_DIAG_DEL_INS, nw_py.GAP_TIE_EXTENDS, nw_py.CASE_SENSITIVE)
    report = {}
    try:
        for combo in itertools.product((True, False), repeat=3):
            nw_py.H_PRIORITY_DIAG_DEL_INS, nw_py.GAP_TIE_EXTENDS, nw_py.CASE_SENSITIVE = combo
            report[combo] = sum(nw_py.nw_traceback(a, b)[:3] != w for (a, b), w in zip(pairs, want))
    finally:
        nw_py.H_PRIORITY_DIAG_DEL_INS, nw_py.GAP_TIE_EXTENDS, nw_py.CASE_SENSITIVE = saved
    return report, saved, len(pairs)


def main():
    parasail = real_parasail()
    if parasail is None:
        print("no real parasail importable here: the tie rules stay UNPINNED (see oracle/nw_oracle.c)")
        return 2
    report, committed, n = compare(parasail)
    print(f"parasail {getattr(parasail, '__version__', '?')}: {n} alignments, traceback strings compared")
    for combo, bad in sorted(report.items(), key=lambda kv: kv[1]):
        tag = "  <- committed rules" if combo == committed else ""
        print(f"  diag>del>ins={combo[0]!s:5} ties_extend={combo[1]!s:5} case_sensitive={combo[2]!s:5}: "
              f"{bad} mismatches{tag}")
    if report[committed] == 0:
        print("PINNED: the committed rules reproduce parasail on every case")
        return 0
    print("the committed rules DIFFER from parasail: flip the switches in oracle/nw_oracle.c, oracle/nw_py.py "
          "and phamclust_b200/csrc/nw_rules.h to the best combination above")
    return 1


if __name__ == "__main__":
    sys.path.insert(0, str(_HERE.parent))
    sys.exit(main())
