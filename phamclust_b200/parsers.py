"""Input parsers with the reference's behaviour
(/root/reference/src/phamclust/scripts/phamclust.py:21-83)."""
import logging
import pathlib

from phamclust_b200.genome import Genome


def load_genomes_from_tsv(filepath):
    """2- or 3-column TSV (name, pham[, translation]); 2 columns -> translation
    "M"; ``rstrip()`` eats a trailing empty third column; repeated (name, pham)
    rows append paralogs in file order; genomes come back in first-seen order."""
    genomes = dict()
    with open(filepath, "r") as handle:
        for line in handle:
            row = line.rstrip().split("\t")
            if len(row) == 3:
                name, pham, translation = row
            elif len(row) == 2:
                (name, pham), translation = row, "M"
            else:
                raise ValueError("input file must either 2 or 3 columns")
            genome = genomes.get(name)
            if genome is None:
                genome = genomes[name] = Genome(name)
            genome.add(pham, translation)
    return list(genomes.values())


def load_genomes_from_fasta_dir(filepath):
    """One FASTA (.fasta/.faa/.fa) per genome, name = file stem, headers
    ``>name=..|pham=..|n=..``; a duplicate stem is a ValueError."""
    filepath = pathlib.Path(filepath)
    genomes = dict()
    for f in filepath.iterdir():
        if f.suffix not in (".fasta", ".faa", ".fa"):
            logging.debug(f"{f} does not appear to be FASTA - skipping")
            continue
        if f.is_dir():
            logging.debug(f"{f} appears to be a directory - skipping")
            continue
        if f.stem in genomes:
            raise ValueError(f"duplicate genome name detected {f.stem}")
        genomes[f.stem] = Genome(f.stem)
        genomes[f.stem].load(f)
    return list(genomes.values())
