#!/usr/bin/env python
"""Golden fixture for row f3 (FASTA ingest): the UNMODIFIED reference parser, src/utils/data_utils.py
DataLoader.parse_sequences (:182-213), run over the edge-case files of tests/fasta_cases.py (build container only:
needs /root/reference; same third-party shim as make_golden.py).  For every case the fixture holds the (id, sequence)
records the reference yields, as one '\\x1f'-joined string each.

Outputs: tests/golden/fasta_cases.npz.    Run: python tests/golden/make_golden_fasta.py
"""
import contextlib
import io
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
from tests.fasta_cases import CASES  # noqa: E402


def main():
    _, ref_du, *_ = mg.import_reference()
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, data in sorted(CASES.items()):
            path = os.path.join(tmp, name + ".fasta")
            with open(path, "wb") as fh:
                fh.write(data)
            with contextlib.redirect_stdout(io.StringIO()):          # the reference prints its parse errors
                recs = list(ref_du.DataLoader.parse_sequences(path))
            out[name + "_ids"] = np.array("\x1f".join(r[0] for r in recs))
            out[name + "_seqs"] = np.array("\x1f".join(r[1] for r in recs))
            out[name + "_count"] = np.array(len(recs))
            print(f"{name}: {len(recs)} records")
    np.savez_compressed(os.path.join(HERE, "fasta_cases.npz"), **out)


if __name__ == "__main__":
    main()
