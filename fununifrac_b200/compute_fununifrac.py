#!/usr/bin/env python
"""CLI of the all-pairs FunUniFrac path — the flags of the reference's ``fununifrac/compute_fununifrac.py:13-35``
(``-e -fd -fp -o -i -f -a -t -b --diffab -v --unweighted --L2 --Ppushed``) plus ``--validate`` (float64 kernels).

    python -m fununifrac_b200.compute_fununifrac -e tree.txt -fd profiles/ -o out/ -b ko00001 -a median_abund
"""
import argparse

from fununifrac_b200 import cli_flags, compute_all_pairwise_fununifrac


def build_parser() -> argparse.ArgumentParser:
    return cli_flags.build(
        "All-pairs functional UniFrac between the samples of a directory of sourmash gather CSVs, over a KEGG "
        "Orthology tree with edge lengths, computed on a B200.",
        ["edge_list", "file_dir", "file_pattern", "out_dir", "out_id", "force", "abundance_key", "threads", "brite",
         "diffab", "verbose", "unweighted", "L2", "Ppushed", "validate", "python_ingest"],
        required=("edge_list", "file_dir", "out_dir"))


def main(argv=None):
    compute_all_pairwise_fununifrac.main(build_parser().parse_args(argv))


if __name__ == '__main__':
    main()
