"""CUDA-graph capture of the DirectGCN inner loop.

The reference trains each n-gram level for up to 500 epochs on ONE fixed graph
(protgram_directgcn_trainer.py:76-108) and then extracts embeddings (:380).  At n-gram-graph sizes
that loop is launch-bound (hundreds of small kernels per epoch), so the whole
`forward -> nll + L2 term -> backward -> Adam -> eval forward` sequence is captured once into a CUDA
graph over static buffers and replayed; a new graph with the same node count only needs its CSR
arrays copied into the static buffers (kernel grids depend on the node count, never on nnz).
"""
from __future__ import annotations

import copy
from typing import Optional

import torch
import torch.nn.functional as F

from .models_utils import EmbeddingProcessor
from .protgram_directgcn import Data, register_symmetric_structure


class GraphedDirectGCNStep:
    """One captured training step (+ optional eval-mode embedding extraction) of ProtGramDirectGCN.

    `optimizer` must be capturable (e.g. torch.optim.Adam(..., capturable=True)).  The loss follows
    the reference trainer: nll_loss(log_probs, y) + l2_lambda * sum ||p||^2 (the L2 term enters as
    its exact gradient 2*l2_lambda*p; its value is added to the reported loss)."""

    def __init__(self, model, optimizer, x: torch.Tensor, labels: torch.Tensor, num_nodes: int, pattern_capacity: int,
                 l2_lambda: float = 0.0, with_extraction: bool = True, warmup: int = 2):
        dev = x.device
        self.model, self.opt, self.l2_lambda, self.with_extraction = model, optimizer, float(l2_lambda), with_extraction
        self.num_nodes, self.capacity = int(num_nodes), int(pattern_capacity)
        self.x, self.labels = x, labels
        self.rowptr = torch.zeros(num_nodes + 1, dtype=torch.int64, device=dev)
        self.col = torch.zeros(self.capacity, dtype=torch.int32, device=dev)
        self.vals = [torch.zeros(self.capacity, dtype=torch.float32, device=dev) for _ in range(3)]
        self._ei = torch.zeros((2, 1), dtype=torch.int64, device=dev)  # placeholder: the CSR is pre-registered (rides on the tensor)
        self._ei._pg_placeholder = True     # get_structure refuses to build a structure from it
        register_symmetric_structure(self._ei, tuple(self.vals), self.num_nodes, self.rowptr, self.col)
        self.data = Data(x=self.x, edge_index_in=self._ei, edge_weight_in=self.vals[0], edge_index_out=self._ei,
                         edge_weight_out=self.vals[1], edge_index_undirected_norm=self._ei,
                         edge_weight_undirected_norm=self.vals[2], num_nodes=self.num_nodes)
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.loss: Optional[torch.Tensor] = None
        self.emb: Optional[torch.Tensor] = None
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.kernels_per_replay = 0
        self.replays = 0
        self._no_long_rows_checked = True  # set False to validate every load_structure (costs a host sync)
        self._warmup = warmup

    def load_structure(self, rowptr: torch.Tensor, col: torch.Tensor, val_in: torch.Tensor, val_out: torch.Tensor,
                       val_und: torch.Tensor) -> None:
        """Copy a (symmetric, shared-pattern) propagation structure into the static buffers."""
        p = int(col.numel())
        if rowptr.numel() != self.num_nodes + 1 or p > self.capacity:
            raise ValueError(f"structure does not fit the captured step (nodes {rowptr.numel() - 1} vs {self.num_nodes}, "
                             f"nnz {p} vs capacity {self.capacity})")
        if self.graph is not None and not self._no_long_rows_checked:
            # the long-row split of the SpMM is planned from rowptr at capture time; a replayed graph
            # must therefore stay in the "no long rows" regime (always true for n-gram graphs: <= 2*sigma+1)
            from .. import _native as nat
            if int((rowptr[1:] - rowptr[:-1]).max()) > nat.SpmmPlan.CHUNK:
                raise ValueError("captured step cannot take a structure with rows longer than the SpMM chunk; re-capture")
        self.rowptr.copy_(rowptr, non_blocking=True)
        self.col[:p].copy_(col, non_blocking=True)
        for dst, src in zip(self.vals, (val_in, val_out, val_und)):
            dst[:p].copy_(src, non_blocking=True)

    def _step(self):
        self.model.train()
        self.opt.zero_grad(set_to_none=True)
        nll = self.model.nll_loss(self.data, self.labels)   # fused decoder output + log_softmax + nll (row f1)
        nll.backward()
        loss = nll.detach()
        if self.l2_lambda > 0.0:
            l2 = torch.stack(torch._foreach_norm(self.params)).square().sum()
            # reference trainer :96: the L2 term covers EVERY trainable parameter, also those the loss did not reach
            # (grad None after zero_grad(set_to_none=True)): their whole gradient is the decay term
            for p in self.params:
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
            torch._foreach_add_([p.grad for p in self.params], self.params, alpha=2.0 * self.l2_lambda)
            loss = loss + self.l2_lambda * l2
        self.opt.step()
        emb = None
        if self.with_extraction:
            # extract_gcn_node_embeddings (models_utils.py:265-273) keeps only the embedding: run the layer
            # stack + L2 normalisation and skip the decoder / log_softmax whose output it discards
            self.model.eval()
            with torch.no_grad():
                emb = EmbeddingProcessor.l2_normalize_torch(self.model.embed(self.data), eps=self.model.l2_eps)
        return loss, emb

    def capture(self) -> None:
        """Warm up on a side stream (state restored afterwards), then capture."""
        saved_model = copy.deepcopy(self.model.state_dict())
        saved_opt = copy.deepcopy(self.opt.state_dict())
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self._warmup):
                self._step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.model.load_state_dict(saved_model)
        if saved_opt["state"]:
            self.opt.load_state_dict(saved_opt)
        else:  # fresh optimizer: reset the moments the warm-up created
            for st in self.opt.state.values():
                for v in st.values():
                    if torch.is_tensor(v):
                        v.zero_()
        self.graph = torch.cuda.CUDAGraph()
        self.opt.zero_grad(set_to_none=True)
        from .. import _native as nat
        before = nat.kernel_launches()
        with torch.cuda.graph(self.graph):
            self.loss, self.emb = self._step()
        self.kernels_per_replay = nat.kernel_launches() - before  # libpgb200 kernels inside the captured graph

    def replay(self):
        if self.graph is None:
            self.capture()
        self.graph.replay()
        self.replays += 1
        return self.loss, self.emb
