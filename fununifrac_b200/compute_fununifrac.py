#!/usr/bin/env python
"""CLI of the all-pairs FunUniFrac path — the flags of the reference's ``fununifrac/compute_fununifrac.py:13-35``
(``-e -fd -fp -o -i -f -a -t -b --diffab -v --unweighted --L2 --Ppushed``) plus ``--validate`` (float64 kernels).

    python -m fununifrac_b200.compute_fununifrac -e tree.txt -fd profiles/ -o out/ -b ko00001 -a median_abund
"""
import argparse
import logging
import multiprocessing

from fununifrac_b200 import compute_all_pairwise_fununifrac


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(description="This script will take a directory of sourmash gather results and "
                                                 "a weighted edge list representing the KEGG hierarchy and compute "
                                                 "all pairwise functional unifrac distances.")
    parser.add_argument('-e', '--edge_list',
                        help='Input edge list file of the KEGG hierarchy. Must have lengths in the '
                             'third column.', required=True)
    parser.add_argument('-fd', '--file_dir', help="Directory of sourmash files.", required=True)
    parser.add_argument('-fp', '--file_pattern', help="Pattern to match files in the directory. Default is "
                                                      "'*_gather.csv'", default='*_gather.csv')
    parser.add_argument('-o', '--out_dir', help='Output directory name.', required=True)
    parser.add_argument('-i', '--out_id', help='Test purpose: give an identifier to the output file so that tester can recognize it')
    parser.add_argument('-f', '--force', help='Overwrite the output file if it exists', action='store_true')
    parser.add_argument('-a', '--abundance_key',
                        help='Key in the gather results to use for abundance. Default is `f_unique_weighted`',
                        default='f_unique_weighted')
    parser.add_argument('-t', '--threads', help='Number of threads to use. Default is half the cores available.',
                        default=int(multiprocessing.cpu_count() / 2), type=int)
    parser.add_argument('-b', '--brite', help='Use the subtree of the KEGG hierarchy rooted at the given BRITE ID. '
                                              'eg. brite:ko00001', default="ko00001", type=str, required=False)
    parser.add_argument('--diffab', action='store_true', help='Also return the difference abundance vectors.')
    parser.add_argument('-v', help="Be verbose", action="store_const", dest="loglevel", const=logging.INFO,
                        default=logging.WARNING)
    parser.add_argument('--unweighted', help="Compute unweighted unifrac instead of the default weighted version",
                        action="store_true")
    parser.add_argument('--L2', help="Use L2 UniFrac instead of L1", action="store_true")
    parser.add_argument('--Ppushed', help="Flag indicating you want the pushed vectors to be saved.", action="store_true")
    parser.add_argument('--validate', help="Run the float64 validation kernels instead of the fp32 production path.",
                        action="store_true")
    parser.add_argument('--python-ingest', dest="python_ingest", action="store_true",
                        help="Read the gather CSVs with pandas and a process pool (the reference's route) instead of "
                             "the native C++ ingest; the profiles are bit-identical either way.")
    return parser


def main(argv=None):
    compute_all_pairwise_fununifrac.main(build_parser().parse_args(argv))


if __name__ == '__main__':
    main()
