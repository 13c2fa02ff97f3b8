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

    @staticmethod
    def save_object(obj: Any, path: str) -> None:
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        with open(path, "wb") as fh:
            pickle.dump(obj, fh, protocol=pickle.HIGHEST_PROTOCOL)

    @staticmethod
    def load_object(path: str) -> Any:
        if not os.path.exists(path):
            raise FileNotFoundError(f"File not found: {path}")
        with open(path, "rb") as fh:
            return pickle.load(fh)
