"""FASTA reading/writing with the reference's semantics
(/root/reference/src/phamclust/fasta.py:4-74): multi-line sequences are
concatenated, every record must start with '>'."""


def read_fasta(filename):
    """Yield ``(header, sequence)`` per record."""
    header, chunks = None, []
    with open(filename, "r") as handle:
        first = True
        if handle.read(1) == "":         # upstream: next() on an empty file inside a generator (fasta.py:17)
            raise RuntimeError("generator raised StopIteration")
        handle.seek(0)
        for line in handle:
            if first and line[0] != ">":
                raise ValueError("records in FASTA files must start with '>'")
            first = False
            if line[0] == ">":
                if header is not None:
                    yield header, "".join(chunks)
                header, chunks = line[1:].rstrip(), []
            else:
                chunks.append(line.rstrip())
    if header is not None:
        yield header, "".join(chunks)


def write_fasta(records, filename, mode="w", wrap=None):
    """Write ``(header, sequence)`` records, optionally wrapping sequence lines."""
    with open(filename, mode) as handle:
        for header, sequence in records:
            handle.write(f">{header}\n")
            if wrap is None or wrap <= 0:
                handle.write(f"{sequence}\n")
            else:
                for i in range(0, len(sequence), wrap):
                    handle.write(sequence[i:i + wrap] + "\n")
    return filename
