"""Existence checks that raise, as the reference's ``data/__init__.py:20-36`` does."""
import glob
import os


def check_data_abspath(path, raise_if_not_found=True):
    if not os.path.exists(path):
        if raise_if_not_found:
            raise Exception(f"data not found: {path} does not exist.")
        return False
    return True


def check_data_abspaths(pattern, raise_if_not_found=True):
    abspaths = glob.glob(pattern)
    if len(abspaths) == 0:
        if raise_if_not_found:
            raise Exception(f"no data found for pattern: {pattern}")
        return False
    return True
