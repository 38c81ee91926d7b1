"""The command-line flags of the two drivers, in one table.

Flag spellings, destinations and defaults are those of the reference CLIs (``fununifrac/compute_fununifrac.py:13-35``,
``fununifrac/compute_fununifrac_combined.py:33-53``) so that existing command lines keep working; the help texts are
this package's own."""
import argparse
import logging
import multiprocessing

_FLAGS = {
    "edge_list": (("-e", "--edge_list"),
                  dict(help="KEGG hierarchy as a tab-separated edge list: parent, child, edge length per line.")),
    "file_dir": (("-fd", "--file_dir"), dict(help="Directory holding the sourmash gather CSVs, one per sample.")),
    "file_pattern": (("-fp", "--file_pattern"),
                     dict(default="*_gather.csv", help="Glob, relative to --file_dir, selecting the sample files "
                                                       "(default: %(default)s).")),
    "file": (("-f", "--file"), dict(help="One CSV with a 'name' column of KO ids and one abundance column per sample.")),
    "out_dir": (("-o", "--out_dir"), dict(help="Directory the result files are written to.")),
    "out_id": (("-i", "--out_id"), dict(help="Identifier used in the result file names (default: a time stamp, or "
                                             "'fununifrac_out' for the combined-table driver).")),
    "force": (("-f", "--force"), dict(action="store_true", help="Write into an existing output location.")),
    "abundance_key": (("-a", "--abundance_key"),
                      dict(default="f_unique_weighted", help="Gather column that carries the abundance "
                                                             "(default: %(default)s).")),
    "threads": (("-t", "--threads"),
                dict(type=int, default=int(multiprocessing.cpu_count() / 2),
                     help="Threads (native ingest) or processes (--python-ingest) reading the sample files "
                          "(default: half the cores).")),
    "brite": (("-b", "--brite"), dict(type=str, default="ko00001", required=False,
                                      help="Root the tree at this BRITE id, e.g. ko00001 (default: %(default)s).")),
    "diffab": (("--diffab",), dict(action="store_true", help="Also produce the per-pair difference vectors.")),
    "verbose": (("-v",), dict(action="store_const", dest="loglevel", const=logging.INFO, default=logging.WARNING,
                              help="Log the milestones of the run.")),
    "unweighted": (("--unweighted",), dict(action="store_true", help="Presence/absence profiles instead of abundances.")),
    "L2": (("--L2",), dict(action="store_true", help="Squared-L2 FunUniFrac (sqrt-length weights) instead of L1.")),
    "Ppushed": (("--Ppushed",), dict(action="store_true", help="Also save the pushed-up profile of every sample.")),
    "validate": (("--validate",), dict(action="store_true",
                                       help="float64 validation kernels (the reference's operation order) instead of the "
                                            "fp32 / int8 production path.")),
    "python_ingest": (("--python-ingest",),
                      dict(dest="python_ingest", action="store_true",
                           help="Read the gather CSVs with pandas and a process pool instead of the native C++ ingest "
                                "(same profiles, bit for bit).")),
}


def build(description: str, names, required=()) -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(description=description)
    for name in names:
        spellings, options = _FLAGS[name]
        options = dict(options)
        if name in required:
            options["required"] = True
        parser.add_argument(*spellings, **options)
    return parser
