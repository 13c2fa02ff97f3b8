"""Host-side corpus packing for hot path A: padded sequences -> the byte buffer the count kernel
reads (include/pgb200.h, "corpus buffer"), the data-defined alphabet, and node-name decoding."""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import numpy as np
import torch

from .. import _native as nat

SEP = 0xFF


def pack_sequences(seqs: Sequence[str], global_first: bool = True) -> np.ndarray:
    """seq_i -> [' ' if global sequence #0] seq_i ' ' 0xFF  (reference data_builder.py:29-35,97-102).
    Raises ValueError for non-ASCII text (the reference windows per code point; the byte kernel
    cannot reproduce that, so it refuses rather than silently diverging)."""
    parts: List[bytes] = []
    for i, s in enumerate(seqs):
        try:
            b = s.encode("ascii")
        except UnicodeEncodeError as exc:
            raise ValueError(f"sequence {i} holds non-ASCII characters; unsupported on the CUDA path") from exc
        parts.append((b" " if (i == 0 and global_first) else b"") + b + b" \xff")
    buf = np.frombuffer(b"".join(parts), dtype=np.uint8)
    return buf


def stream_chunks(seqs: Iterable[str], chunk_bytes: int = 1 << 28, rank: int = 0, world: int = 1, block: int = 4096):
    """Pack an iterator of sequences into corpus-buffer chunks of ~chunk_bytes, cut at sequence
    boundaries, without ever holding the whole corpus as Python strings.  With world > 1 the
    sequences are dealt to ranks in blocks of `block` (counting is order independent); the leading
    space of global sequence #0 stays with the rank that owns it.  Yields np.uint8 arrays."""
    parts: List[bytes] = []
    size = 0
    for i, s in enumerate(seqs):
        if (i // block) % world != rank:
            continue
        try:
            b = s.encode("ascii")
        except UnicodeEncodeError as exc:
            raise ValueError(f"sequence {i} holds non-ASCII characters; unsupported on the CUDA path") from exc
        piece = (b" " if i == 0 else b"") + b + b" \xff"
        parts.append(piece)
        size += len(piece)
        if size >= chunk_bytes:
            yield np.frombuffer(b"".join(parts), dtype=np.uint8)
            parts, size = [], 0
    if parts:
        yield np.frombuffer(b"".join(parts), dtype=np.uint8)


class NonAsciiSequence(ValueError):
    pass


def stream_chunks_native(fasta_path: str, chunk_bytes: int = 1 << 28, rank: int = 0, world: int = 1, block: int = 4096,
                         pinned: bool = False, stats: dict = None):
    """FASTA file -> corpus-buffer chunks without a Python string per sequence (row f3 of SURVEY.md 8): the native
    reader (csrc/fasta.cu) applies the record rules of DataLoader.parse_sequences and the padding rule of
    pack_sequences while copying out of the mmap'ed file.  Same chunks as
    `stream_chunks((s for _, s in DataLoader.parse_sequences(path)), ...)` up to where the cuts fall.
    Yields uint8 torch tensors (pinned when asked: they go to the GPU with one async copy).
    stats (optional dict) receives 'sequences' (all ranks) and 'stopped_early'."""
    import ctypes
    import os
    lib = nat.load()
    reader = lib.pg_fasta_open(os.path.normpath(fasta_path).encode())
    if not reader:
        raise FileNotFoundError(lib.pg_last_error().decode())
    try:
        # a packed record never takes more bytes than its text (header >= 2 bytes pay for the ' ' 0xFF): file size + 1 bounds it
        cap = max(min(int(chunk_bytes), int(lib.pg_fasta_file_bytes(reader)) + 64), 1 << 12)
        while True:
            buf = torch.empty(cap, dtype=torch.uint8, pin_memory=bool(pinned and torch.cuda.is_available()))
            got = int(lib.pg_fasta_next_chunk(reader, ctypes.c_void_p(buf.data_ptr()), cap, rank, world, block))
            if got == nat.PG_FASTA_ETOOSMALL:
                cap *= 4          # one record longer than the chunk: grow and retry (the reader did not advance)
                continue
            if got == nat.PG_FASTA_ENONASCII:
                raise NonAsciiSequence(lib.pg_last_error().decode())
            if got < 0:
                raise nat.NativeError(f"pg_fasta_next_chunk failed ({got}): {lib.pg_last_error().decode()}")
            if got == 0:
                break
            yield buf[:got]
        if stats is not None:
            stats["sequences"] = int(lib.pg_fasta_records(reader))
            stats["stopped_early"] = bool(lib.pg_fasta_stopped_early(reader))
    finally:
        lib.pg_fasta_close(reader)


class PackedChunk:
    """A corpus chunk kept in HOST memory in the 5-bit format (8 symbols in 5 bytes, include/pgb200.h): what crosses
    PCIe when a corpus is streamed from the host (chunks beyond the HBM budget are uploaded once per n level)."""

    def __init__(self, packed: torch.Tensor, n_symbols: int):
        self.packed, self.n_symbols = packed, int(n_symbols)

    @property
    def size(self) -> int:               # logical corpus bytes
        return self.n_symbols


def pack5(host_buf, pinned: bool = True):
    """Corpus buffer (host bytes) -> PackedChunk, or None when some byte has no 5-bit code (keep the byte format)."""
    import ctypes
    host = torch.from_numpy(np.ascontiguousarray(host_buf)) if not torch.is_tensor(host_buf) else host_buf.contiguous()
    n = int(host.numel())
    lib = nat.load()
    out = torch.empty(int(lib.pg_pack5_bytes(n)), dtype=torch.uint8, pin_memory=bool(pinned and torch.cuda.is_available()))
    got = int(lib.pg_pack5_host(ctypes.c_void_p(host.data_ptr()), n, ctypes.c_void_p(out.data_ptr())))
    if got == nat.PG_EPACK:
        return None
    if got < 0:
        raise nat.NativeError(f"pg_pack5_host failed ({got}): {lib.pg_last_error().decode()}")
    return PackedChunk(out[:got], n)


def unpack5(d_packed: torch.Tensor, n_symbols: int, out: torch.Tensor = None) -> torch.Tensor:
    """Packed bytes on the device -> the corpus buffer the kernels read (a view of `out` when it is large enough)."""
    if out is None or out.numel() < n_symbols:
        out = torch.empty(max(int(n_symbols), 16), dtype=torch.uint8, device=d_packed.device)
    nat.call("pg_unpack5", nat.ptr(d_packed), int(n_symbols), nat.ptr(out), nat.stream_ptr())
    return out[:n_symbols]


def chunk_to_device(c, device) -> torch.Tensor:
    """Device tensor / host byte array / PackedChunk -> corpus buffer on `device`."""
    if isinstance(c, PackedChunk):
        return unpack5(to_device(c.packed, device), c.n_symbols)
    if torch.is_tensor(c) and c.is_cuda:
        return c
    return to_device(c, device)


def read_fasta_parallel(fasta_path: str, threads: int = 0, rank: int = 0, world: int = 1, block: int = 4096,
                        pinned: bool = False, stats: dict = None) -> torch.Tensor:
    """Whole FASTA file -> ONE corpus buffer (this rank's records), parsed by several host threads over file ranges cut at
    header lines (csrc/fasta.cu:pg_fasta_pack_parallel).  Same bytes as concatenating stream_chunks_native()."""
    import ctypes
    import os
    lib = nat.load()
    path = os.path.normpath(fasta_path)
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    threads = int(threads) if threads else max(1, min(32, len(os.sched_getaffinity(0))))
    cap = os.path.getsize(path) + 64
    buf = torch.empty(cap, dtype=torch.uint8, pin_memory=bool(pinned and torch.cuda.is_available()))
    n_rec, stopped = ctypes.c_int64(0), ctypes.c_int(0)
    got = int(lib.pg_fasta_pack_parallel(path.encode(), ctypes.c_void_p(buf.data_ptr()), cap, threads, rank, world, block,
                                         ctypes.byref(n_rec), ctypes.byref(stopped)))
    if got == nat.PG_FASTA_ENONASCII:
        raise NonAsciiSequence(lib.pg_last_error().decode())
    if got < 0:
        raise nat.NativeError(f"pg_fasta_pack_parallel failed ({got}): {lib.pg_last_error().decode()}")
    if stats is not None:
        stats["sequences"], stats["stopped_early"] = int(n_rec.value), bool(stopped.value)
    return buf[:got]


def stream_fasta_windows(fasta_path: str, window_bytes: int = 1 << 30, threads: int = 0, rank: int = 0, world: int = 1,
                         block: int = 4096, pinned: bool = False, stats: dict = None):
    """Files that do not fit host memory as one corpus buffer: walk the file in windows of about window_bytes (cut at header
    lines), each parsed by several host threads (csrc/fasta.cu:pg_fasta_pack_window).  Yields one uint8 tensor per window
    (this rank's records; empty windows are skipped); concatenated they are the bytes of read_fasta_parallel()."""
    import ctypes
    import os
    lib = nat.load()
    path = os.path.normpath(fasta_path)
    reader = lib.pg_fasta_open(path.encode())
    if not reader:
        raise FileNotFoundError(lib.pg_last_error().decode())
    threads = int(threads) if threads else max(1, min(32, len(os.sched_getaffinity(0))))
    try:
        size = int(lib.pg_fasta_file_bytes(reader))
        pos, first, n_rec, stopped = ctypes.c_int64(0), 0, ctypes.c_int64(0), ctypes.c_int(0)
        window = max(int(window_bytes), 1)
        cap = min(window, size) + (1 << 16)
        while pos.value < size and not stopped.value:
            buf = torch.empty(cap, dtype=torch.uint8, pin_memory=bool(pinned and torch.cuda.is_available()))
            got = int(lib.pg_fasta_pack_window(reader, ctypes.byref(pos), window, first, ctypes.c_void_p(buf.data_ptr()), cap, threads,
                                               rank, world, block, ctypes.byref(n_rec), ctypes.byref(stopped)))
            if got == nat.PG_FASTA_ETOOSMALL:
                cap *= 2          # a record running far past the nominal window end: grow and retry (pos did not move)
                continue
            if got == nat.PG_FASTA_ENONASCII:
                raise NonAsciiSequence(lib.pg_last_error().decode())
            if got < 0:
                raise nat.NativeError(f"pg_fasta_pack_window failed ({got}): {lib.pg_last_error().decode()}")
            first += int(n_rec.value)
            if got:
                yield buf[:got]
        if stats is not None:
            stats["sequences"], stats["stopped_early"] = first, bool(stopped.value)
    finally:
        lib.pg_fasta_close(reader)


def split_at_separators(buf: torch.Tensor, chunk_bytes: int):
    """Views of a host corpus buffer of about chunk_bytes each, cut right after a sequence separator."""
    n = int(buf.numel())
    out, lo = [], 0
    arr = buf.numpy()
    while lo < n:
        hi = min(n, lo + max(int(chunk_bytes), 1))
        if hi < n:
            back = np.flatnonzero(arr[lo:hi] == SEP)
            if back.size:
                hi = lo + int(back[-1]) + 1
            else:                                   # one sequence longer than the chunk: extend to its separator
                fwd = np.flatnonzero(arr[hi:] == SEP)
                hi = hi + int(fwd[0]) + 1 if fwd.size else n
        out.append(buf[lo:hi])
        lo = hi
    return out


def to_device(buf: np.ndarray, device) -> torch.Tensor:
    """Pinned staging + async H2D of the corpus buffer (16 B aligned by the allocator)."""
    host = torch.from_numpy(np.ascontiguousarray(buf).copy()) if not isinstance(buf, torch.Tensor) else buf
    if torch.device(device).type != "cuda":
        return host
    if not host.is_pinned():
        host = host.pin_memory()
    return host.to(device, non_blocking=True)


class CorpusUploader:
    """Double-buffered pinned-host -> HBM upload of corpus buffers on a side stream, so the copy of
    buffer i+1 runs under the kernels that consume buffer i (PCIe Gen5 moves ~55 GB/s, the count
    kernel consumes > 500 GB/s: an un-overlapped upload is the end-to-end bottleneck of hot path A).

        up = CorpusUploader(device)
        up.submit(host_chunk_0)
        for i in range(n):
            if i + 1 < n: up.submit(host_chunk_{i+1})      # prefetch
            d_buf = up.acquire()                            # current stream waits for the copy of chunk i
            ... launch kernels reading d_buf ...
            up.release()                                    # slot reusable once those kernels are done
    """

    def __init__(self, device, slots: int = 2):
        nat.require_cuda()
        self.device = torch.device(device)
        # only the host-logic tests (native entry points swapped for their CPU specification) get here without CUDA
        self.passthrough = self.device.type != "cuda"
        self.copy_stream = None if self.passthrough else torch.cuda.Stream(device=self.device)
        self.slots = [None] * slots                     # device buffers (grown on demand)
        self.free_ev = [None] * slots                   # consumer-done events
        self.pending = []                               # (slot, view, copy-done event) in submit order
        self.in_use = []                                # acquired, not yet released
        self._next = 0

    def submit(self, host_buf) -> None:
        host = torch.from_numpy(np.ascontiguousarray(host_buf)) if not torch.is_tensor(host_buf) else host_buf
        if self.passthrough:
            self.pending.append((0, host, None, host))
            return
        if not host.is_pinned():
            host = host.pin_memory()
        k = self._next
        self._next = (k + 1) % len(self.slots)
        if any(p[0] == k for p in self.pending) or k in self.in_use:
            raise RuntimeError("CorpusUploader: all slots are busy (acquire/release before submitting more)")
        n = host.numel()
        with torch.cuda.stream(self.copy_stream):
            if self.slots[k] is None or self.slots[k].numel() < n:
                self.slots[k] = torch.empty(max(n, 16), dtype=torch.uint8, device=self.device)
            if self.free_ev[k] is not None:
                self.copy_stream.wait_event(self.free_ev[k])
            view = self.slots[k][:n]
            view.copy_(host, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.pending.append((k, view, ev, host))        # `host` kept alive until the copy is consumed

    def acquire(self) -> torch.Tensor:
        k, view, ev, _host = self.pending.pop(0)
        if self.passthrough:
            return view
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        view.record_stream(cur)   # the slot was allocated on the copy stream: tell the allocator who else reads it
        self.in_use.append(k)
        return view

    def release(self) -> None:
        if self.passthrough:
            return
        k = self.in_use.pop(0)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.free_ev[k] = ev


def discover_alphabet(d_buf, group=None) -> Tuple[np.ndarray, torch.Tensor]:
    """-> (symbols uint8[sigma] ascending, rank_of_byte uint8[256] on the device).
    d_buf: one corpus buffer or a list of chunks (device tensors, or host arrays uploaded one at a time).  With a process group the 256-entry presence table
    is OR-reduced first so every rank packs with the same alphabet."""
    nat.require_cuda()
    chunks = list(d_buf) if isinstance(d_buf, (list, tuple)) else [d_buf]
    dev = next((c.device for c in chunks if torch.is_tensor(c) and c.is_cuda), None) or nat.current_device()
    pres = torch.zeros(256, dtype=torch.int32, device=dev)
    for c in chunks:
        c_dev = chunk_to_device(c, dev)                                         # host chunks stream through one at a time
        nat.call("pg_byte_presence", nat.ptr(c_dev), c_dev.numel(), nat.ptr(pres), nat.stream_ptr())
    if group is not None:
        import torch.distributed as dist
        dist.all_reduce(pres, op=dist.ReduceOp.MAX, group=group)
    present = pres.cpu().numpy() != 0
    if present[128:].any():
        # the kernels treat every byte >= 0x80 as the sequence separator (7-bit ASCII contract)
        raise ValueError("corpus buffer holds non-ASCII bytes; unsupported on the CUDA path")
    return alphabet_from_presence(present, dev)


def alphabet_from_presence(present: np.ndarray, device) -> Tuple[np.ndarray, torch.Tensor]:
    symbols = np.nonzero(present)[0].astype(np.uint8)
    rank = np.zeros(256, dtype=np.uint8)
    rank[symbols] = np.arange(symbols.size, dtype=np.uint8)
    return symbols, torch.from_numpy(rank).to(device)


def decode_nodes(node_code: np.ndarray, symbols: np.ndarray, n: int) -> List[str]:
    """base-sigma codes -> n-gram strings (most significant digit = first character)."""
    sigma = int(symbols.size)
    code = np.asarray(node_code, dtype=np.int64).copy()
    chars = np.empty((code.size, n), dtype=np.uint8)
    for k in range(n - 1, -1, -1):
        chars[:, k] = symbols[code % sigma]
        code //= sigma
    if code.size == 0:
        return []
    # 7-bit ASCII bytes widened to UCS4 code points and viewed as fixed-width unicode: one vectorised
    # cast instead of 8k Python-level decodes (numpy 'U' strips trailing NULs only; NUL never ranks).
    if 0 in symbols:
        return [row.tobytes().decode("ascii") for row in chars]
    return chars.astype(np.uint32).view(f"U{n}").ravel().tolist()


_PINNED_POOL: List[torch.Tensor] = []   # reusable pinned staging buffers (cudaHostAlloc costs ~1 ms a call)


class LazyNodeNames:
    """Node names of one level, decoded lazily: the packed codes start an async device->pinned-host
    copy at construction and are turned into strings on `resolve()` (first access of
    `graph.node_sequences`), so the host-side decode runs under whatever GPU work follows the
    extraction instead of stalling the stream."""

    def __init__(self, node_code: torch.Tensor, symbols: np.ndarray, n: int):
        self.symbols, self.n, self.count = symbols, int(n), int(node_code.numel())
        self._names = None
        if not node_code.is_cuda:
            self._host, self._event = node_code, None
            return
        need = max(self.count, 1)
        buf = next((b for b in _PINNED_POOL if b.numel() >= need), None)
        if buf is not None:
            _PINNED_POOL.remove(buf)
        else:
            buf = torch.empty(max(need, 1 << 16), dtype=torch.int64).pin_memory()
        self._buf = buf
        self._host = buf[:self.count]
        self._host.copy_(node_code, non_blocking=True)
        self._event = torch.cuda.Event()
        self._event.record()

    def __len__(self) -> int:
        return self.count

    def resolve(self) -> List[str]:
        if self._names is None:
            if self._event is not None:
                self._event.synchronize()
            self._names = decode_nodes(self._host.numpy(), self.symbols, self.n)
            if self._event is not None:
                _PINNED_POOL.append(self._buf)
                self._buf = self._host = None
        return self._names
