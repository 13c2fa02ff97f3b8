"""FASTA ingest and object hand-off -- mirrors the two pieces of the reference's
src/utils/data_utils.py that sit on hot path A: DataLoader.parse_sequences (:182-213) and
DataUtils.save_object / load_object (:406-431)."""
from __future__ import annotations

import os
import pickle
from typing import Any, Iterator, List, Optional, Tuple


class DataLoader:
    @staticmethod
    def parse_sequences(fasta_filepath: str) -> Iterator[Tuple[str, str]]:
        """Yield (protein_id, SEQUENCE).  Same record rules as the reference: lines are stripped,
        blank lines skipped, sequence text upper-cased, id = 2nd '|' field if present else the
        first header token, records without sequence text dropped.  A missing file prints an
        error and yields nothing (the reference never raises from here)."""
        fasta_filepath = os.path.normpath(fasta_filepath)
        pid: Optional[str] = None
        chunks: List[str] = []
        try:
            with open(fasta_filepath, "r", encoding="utf-8", errors="ignore") as handle:
                for raw in handle:
                    text = raw.strip()
                    if not text:
                        continue
                    if text.startswith(">"):
                        if pid and chunks:
                            yield pid, "".join(chunks)
                        head = text[1:]
                        fields = head.split("|")
                        pid = fields[1] if len(fields) > 1 and fields[1] else head.split()[0]
                        chunks = []
                    elif pid is not None:
                        chunks.append(text.upper())
            if pid and chunks:
                yield pid, "".join(chunks)
        except FileNotFoundError:
            print(f"Error: FASTA file not found at {fasta_filepath}")
        except Exception as exc:  # noqa: BLE001 - reference prints and continues
            print(f"Error parsing FASTA file {fasta_filepath}: {exc}")


class DataUtils:
    @staticmethod
    def print_header(title: str) -> None:
        bar = "=" * 80
        print(f"\n{bar}\n### {title} ###\n{bar}\n")

    SIDECAR_SUFFIX = ".csr.npz"

    @staticmethod
    def save_object(obj: Any, path: str) -> None:
        """Reference :406-418 (pickle).  A DirectedNgramGraph additionally gets `<path>.csr.npz`: its three propagation
        matrices as ONE shared-pattern CSR (rowptr int64, col int32, three fp32 value arrays: 16 B per stored entry instead
        of the pickle's 3 x 20 B of COO).  The pickle itself is unchanged and stays loadable by the reference; the sidecar
        is optional on the way back in."""
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        with open(path, "wb") as fh:
            pickle.dump(obj, fh, protocol=pickle.HIGHEST_PROTOCOL)
        try:
            csr = obj.propagation_csr_host() if hasattr(obj, "propagation_csr_host") else None
            if csr is not None:
                import numpy as np
                with open(path + DataUtils.SIDECAR_SUFFIX, "wb") as fh:
                    np.savez(fh, **csr)
            elif os.path.exists(path + DataUtils.SIDECAR_SUFFIX):
                os.remove(path + DataUtils.SIDECAR_SUFFIX)      # never leave a stale sidecar next to a new pickle
        except Exception as exc:  # noqa: BLE001 - the sidecar is an optimisation; the pickle is the contract
            print(f"Warning: could not write the CSR sidecar for {path}: {exc}")

    @staticmethod
    def load_object(path: str) -> Any:
        if not os.path.exists(path):
            raise FileNotFoundError(f"File not found: {path}")
        with open(path, "rb") as fh:
            obj = pickle.load(fh)
        side = path + DataUtils.SIDECAR_SUFFIX
        if hasattr(obj, "attach_propagation_csr") and os.path.exists(side) and os.path.getmtime(side) >= os.path.getmtime(path):
            try:
                import numpy as np
                with np.load(side, allow_pickle=False) as z:
                    obj.attach_propagation_csr({k: z[k] for k in z.files})
            except Exception as exc:  # noqa: BLE001
                print(f"Warning: ignoring the CSR sidecar of {path}: {exc}")
        return obj
