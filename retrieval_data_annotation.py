#!/usr/bin/env python
"""CLI shim with the reference's name and arguments:

    python retrieval_data_annotation.py <dataset> <timestep> <threshold>

Runs the B200 engine (rag4dyg_b200.retrieval_data_annotation); reads/writes the same CWD-relative files as
/root/reference/retrieval_data_annotation.py:117-144.
"""
from rag4dyg_b200.retrieval_data_annotation import *  # noqa: F401,F403  (same importable names as the reference)
from rag4dyg_b200.retrieval_data_annotation import main

if __name__ == "__main__":
    main()
