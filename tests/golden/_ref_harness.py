"""Kept for ``make_golden.py`` and older scripts: the import harness lives in ``oracle/reference_import.py``."""
from oracle.reference_import import _AnnData, import_reference, make_adata, ref_run, reference_root  # noqa: F401

REFERENCE_ROOT = "/root/reference"
