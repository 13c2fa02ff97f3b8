"""CPU oracle for hot path B (DirectGCN propagation, fwd + bwd via autograd) -- TEST
INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py cpu_baseline / --impl reference).

Fresh torch-CPU fp32 restatement of (citations into /root/reference):

  propagate              <- PyG MessagePassing(aggr='add'), default flow source_to_target, as
                            called at src/models/protgram_directgcn.py:101-112 with
                            message() = :137-140:   out[ei[1]] += w * x[ei[0]]
  directgcn_layer        <- protgram_directgcn.py:93-135 (6 propagates, 6 biases, 5 gates, constant)
  apply_pe               <- :182-193
  l2_normalize           <- src/utils/models_utils.py:139-147
  protgram_forward       <- :195-222 (PE, [layer + res_proj -> leaky_relu -> dropout]*L,
                            decoder Linear-ReLU-Dropout-Linear, log_softmax, l2-normalised emb)

Parameters are passed as a plain dict with the reference's state_dict key names
("convs.0.lin_main_in.weight", "convs.0.C_in_vec", "res_projs.0.weight", "pe_layer.weight",
"decoder_fc.0.weight", ...), so a reference checkpoint drives it unchanged.

Parity status: PINNED against tests/golden/model_*.npz (reference-generated outputs, per-layer
activations and gradients).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


def propagate(ei: torch.Tensor, x: torch.Tensor, ew: Optional[torch.Tensor]) -> torch.Tensor:
    msg = x[ei[0]]
    if ew is not None:
        msg = ew.view(-1, 1) * msg
    return torch.zeros_like(x).index_add_(0, ei[1], msg)


def directgcn_layer(p: Dict[str, torch.Tensor], prefix: str, x, ei_in, ew_in, ei_out, ew_out,
                    ei_un, ew_un, original_indices=None):
    g = lambda k: p[prefix + k]
    lin = lambda k: x @ g(k + ".weight").t()
    shared = lin("lin_shared")
    ic = (propagate(ei_in, lin("lin_main_in"), ew_in) + g("bias_main_in")) + \
         (propagate(ei_in, shared, ew_in) + g("bias_directed_shared_in"))
    oc = (propagate(ei_out, lin("lin_main_out"), ew_out) + g("bias_main_out")) + \
         (propagate(ei_out, shared, ew_out) + g("bias_directed_shared_out"))
    uc = (propagate(ei_un, lin("lin_undirected"), ew_un) + g("bias_undirected")) + \
         (propagate(ei_un, shared, ew_un) + g("bias_undirected_shared"))
    if prefix + "C_in_vec" in p:
        names = ["C_in_vec", "C_out_vec", "C_directed_vec", "C_undirected_vec", "C_all_vec"]
        gates = [g(k) for k in names]
        const = p.get(prefix + "constant")
        if original_indices is not None:
            gates = [t[original_indices] for t in gates]
            const = const[original_indices] if const is not None else None
        if const is None:
            const = 0
    else:
        gates = [g(k) for k in ["C_in", "C_out", "C_directed", "C_undirected", "C_all"]]
        const = 0
    c_in, c_out, c_dir, c_un, c_all = gates
    directed = c_dir * (c_in * ic + c_out * oc)
    return c_all * (c_un * uc + directed) + const


def apply_pe(p, x, n_gram_len: int, one_gram_dim: int):
    w = p.get("pe_layer.weight")
    if w is None:
        return x
    if n_gram_len > 0 and one_gram_dim > 0 and x.shape[1] == n_gram_len * one_gram_dim:
        k = min(n_gram_len, w.shape[0])
        if k > 0:
            xr = x.clone().view(-1, n_gram_len, one_gram_dim)
            xr[:, :k, :] = xr[:, :k, :] + w[:k].unsqueeze(0)
            return xr.view(-1, n_gram_len * one_gram_dim)
    return x


def l2_normalize(h, eps: float = 1e-12):
    return h / (torch.norm(h, p=2, dim=1, keepdim=True) + eps)


def protgram_forward(p: Dict[str, torch.Tensor], x, ei_in, ew_in, ei_out, ew_out, ei_un, ew_un,
                     n_gram_len: int, one_gram_dim: int, original_indices=None,
                     l2_eps: float = 1e-12, return_layers: bool = False):
    """Eval-mode forward (dropout inactive).  -> (log_probs, emb[, per-layer conv outputs])."""
    n_layers = 1 + max(int(k.split(".")[1]) for k in p if k.startswith("convs."))
    h = apply_pe(p, x, n_gram_len, one_gram_dim)
    conv_outs: List[torch.Tensor] = []
    for i in range(n_layers):
        conv = directgcn_layer(p, f"convs.{i}.", h, ei_in, ew_in, ei_out, ew_out, ei_un, ew_un,
                               original_indices)
        conv_outs.append(conv)
        if f"res_projs.{i}.weight" in p:
            res = h @ p[f"res_projs.{i}.weight"].t() + p[f"res_projs.{i}.bias"]
        else:
            res = h
        h = F.leaky_relu(conv + res)
    z = F.relu(h @ p["decoder_fc.0.weight"].t() + p["decoder_fc.0.bias"])
    logits = z @ p["decoder_fc.3.weight"].t() + p["decoder_fc.3.bias"]
    out = (F.log_softmax(logits, dim=-1), l2_normalize(h, l2_eps))
    return out + (conv_outs,) if return_layers else out


def init_params(layer_dims, num_nodes: int, num_classes: int, one_gram_dim: int = 0, max_pe_len: int = 0,
                seed: int = 42) -> Dict[str, torch.Tensor]:
    """Fresh parameters with the reference's shapes and initialisers (protgram_directgcn.py:34-91,
    :156-180): xavier-uniform Linear weights and `constant`, zero biases, unit gates, default
    nn.Linear init for res_projs / decoder.  Leaf tensors with requires_grad=True."""
    g = torch.Generator().manual_seed(seed)
    p: Dict[str, torch.Tensor] = {}

    def xavier(*shape):
        fan_out, fan_in = shape[0], shape[1]
        a = (6.0 / (fan_in + fan_out)) ** 0.5
        return (torch.rand(*shape, generator=g) * 2 - 1) * a

    def linear(prefix, fin, fout):
        bound = 1.0 / fin ** 0.5
        p[prefix + ".weight"] = (torch.rand(fout, fin, generator=g) * 2 - 1) * bound
        p[prefix + ".bias"] = (torch.rand(fout, generator=g) * 2 - 1) * bound

    if one_gram_dim > 0 and max_pe_len > 0:
        p["pe_layer.weight"] = torch.randn(max_pe_len, one_gram_dim, generator=g)
    for i in range(len(layer_dims) - 1):
        fin, fout = layer_dims[i], layer_dims[i + 1]
        pre = f"convs.{i}."
        for k in ("lin_main_in", "lin_main_out", "lin_undirected", "lin_shared"):
            p[pre + k + ".weight"] = xavier(fout, fin)
        for k in ("bias_main_in", "bias_main_out", "bias_undirected", "bias_directed_shared_in",
                  "bias_directed_shared_out", "bias_undirected_shared"):
            p[pre + k] = torch.zeros(fout)
        for k in ("C_in_vec", "C_out_vec", "C_directed_vec", "C_undirected_vec", "C_all_vec"):
            p[pre + k] = torch.ones(num_nodes, 1)
        p[pre + "constant"] = xavier(num_nodes, fout)
        if fin != fout:
            linear(f"res_projs.{i}", fin, fout)
    final = layer_dims[-1]
    hidden = final // 2 if final > 1 else 1
    linear("decoder_fc.0", final, hidden)
    linear("decoder_fc.3", hidden, num_classes)
    return {k: v.requires_grad_(True) for k, v in p.items()}
