"""The drop-in CLI surface: flag spellings, destinations and defaults of the reference's
``fununifrac/compute_fununifrac.py:13-35`` and ``fununifrac/compute_fununifrac_combined.py:33-53``.  CPU only."""
import logging
import multiprocessing

import pytest

from fununifrac_b200 import compute_fununifrac, compute_fununifrac_combined


def test_main_cli_defaults_and_spellings():
    p = compute_fununifrac.build_parser()
    a = p.parse_args(["-e", "tree.txt", "-fd", "profiles", "-o", "out"])
    assert (a.edge_list, a.file_dir, a.out_dir) == ("tree.txt", "profiles", "out")
    assert a.file_pattern == "*_gather.csv" and a.abundance_key == "f_unique_weighted" and a.brite == "ko00001"
    assert a.threads == int(multiprocessing.cpu_count() / 2) and a.out_id is None
    assert a.loglevel == logging.WARNING
    assert not (a.force or a.diffab or a.unweighted or a.L2 or a.Ppushed or a.validate or a.python_ingest)
    b = p.parse_args(["--edge_list", "t", "--file_dir", "d", "--file_pattern", "*.csv", "--out_dir", "o", "--out_id", "x",
                      "--force", "--abundance_key", "median_abund", "--threads", "3", "--brite", "ko00002", "--diffab",
                      "-v", "--unweighted", "--L2", "--Ppushed"])
    assert (b.file_pattern, b.out_id, b.abundance_key, b.threads, b.brite) == ("*.csv", "x", "median_abund", 3, "ko00002")
    assert b.force and b.diffab and b.unweighted and b.L2 and b.Ppushed and b.loglevel == logging.INFO
    c = p.parse_args(["-e", "t", "-fd", "d", "-o", "o", "-fp", "p", "-i", "id", "-f", "-a", "k", "-t", "1", "-b", "b"])
    assert (c.file_pattern, c.out_id, c.force, c.abundance_key, c.threads, c.brite) == ("p", "id", True, "k", 1, "b")
    for missing in (["-fd", "d", "-o", "o"], ["-e", "t", "-o", "o"], ["-e", "t", "-fd", "d"]):
        with pytest.raises(SystemExit):
            p.parse_args(missing)                      # -e, -fd and -o are required, as in the reference


def test_combined_cli_defaults_and_spellings():
    p = compute_fununifrac_combined.build_parser()
    a = p.parse_args([])                               # the reference marks nothing as required here
    assert a.edge_list is None and a.file is None and a.out_dir is None and a.out_id is None
    assert a.abundance_key == "f_unique_weighted" and a.brite == "ko00001" and a.loglevel == logging.WARNING
    assert not (a.diffab or a.unweighted or a.L2 or a.Ppushed or a.validate)
    b = p.parse_args(["-e", "t", "-f", "table.csv", "-o", "o", "-i", "x", "-a", "k", "-b", "ko00002", "--diffab", "-v",
                      "--unweighted", "--L2", "--Ppushed"])
    assert (b.edge_list, b.file, b.out_dir, b.out_id, b.abundance_key, b.brite) == ("t", "table.csv", "o", "x", "k", "ko00002")
    assert b.diffab and b.unweighted and b.L2 and b.Ppushed and b.loglevel == logging.INFO
