"""The reference's sub-blocks called ON THEIR OWN with the reference's signatures -- GCN.forward(graph_em, edge,
210, 160, 280) (gnn_transformer.py:74), Attention.forward (:137), FeedForward.forward (:170), Combination.forward
(:192), CombinationLayer.forward (combination_layer.py:7) -- against the CPU oracle's restatement of the same
block, outputs and every gradient (fp32 parity mode, 1e-4)."""
import numpy as np
import pytest
import torch

from fira_testlib import golden_batch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _sd(module, prefix):
    return {f"{prefix}.{k}": v.detach().cpu().double() for k, v in module.state_dict().items()}


def _close(a, b, tol=1e-4, what=""):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = (a - b).abs().max().item()
    # + 1e-6 absolute: gradients that vanish in exact arithmetic (fc_k.bias under the softmax's shift invariance, the
    # input of a LayerNorm under a constant upstream) are fp32 round-off noise on both sides
    assert err <= tol * b.abs().max().item() + 1e-6, (what, err, b.abs().max().item())


def _grads_match(module, sd_ref, prefix, tol=2e-4):
    for name, p in module.named_parameters():
        ref = sd_ref[f"{prefix}.{name}"].grad
        if ref is None:
            assert p.grad is None or p.grad.abs().max().item() == 0, name
            continue
        _close(p.grad, ref, tol, name)


def _leafify(sd):
    return {k: v.clone().requires_grad_(True) for k, v in sd.items()}


def test_gcn_forward_signature_and_gradients():
    import gnn_transformer as G          # the drop-in shim at the repository root
    import fira_oracle as O
    torch.manual_seed(1)
    m = G.GCN(256, dropout_rate=0.2).to(DEV).eval()
    B = 3
    edge = golden_batch(0, B)[5]                                  # float64 [B,650,650], as the reference feeds it
    x = torch.randn(B, 650, 256)
    xg = x.to(DEV).requires_grad_(True)
    code, sub, ast = m(xg, edge.to(DEV), 210, 160, 280)
    assert code.shape == (B, 210, 256) and sub.shape == (B, 160, 256) and ast.shape == (B, 280, 256)
    sd = _leafify(_sd(m, "g"))
    xr = x.double().requires_grad_(True)
    ref = O.gcn(sd, "g", xr, edge, 0.0, False)
    out = torch.cat((code, sub, ast), 1)
    _close(out, ref, 1e-4, "gcn out")
    w = torch.randn(B, 650, 256)
    (out * w.to(DEV)).sum().backward()
    (ref * w.double()).sum().backward()
    _close(xg.grad, xr.grad, 2e-4, "gcn dx")
    _grads_match(m, sd, "g")


@pytest.mark.parametrize("kind", ["self", "cross"])
def test_attention_forward_signature_and_gradients(kind):
    import gnn_transformer as G
    import fira_oracle as O
    torch.manual_seed(2)
    m = G.Attention(256, 8).to(DEV).eval()
    B, T, S = 4, 30, 370
    q = torch.randn(B, T, 256)
    if kind == "self":
        mem = q
        pad = torch.rand(B, T) > 0.3
        pad[:, 0] = True
        mask = pad[:, None, None, :] & torch.tril(torch.ones(T, T, dtype=torch.bool))[None, None]   # :117
    else:
        mem = torch.randn(B, S, 256)
        mask = torch.rand(B, S) > 0.5                                                                # :120
        mask[:, 0] = True
    qg = q.to(DEV).requires_grad_(True)
    memg = qg if kind == "self" else mem.to(DEV).requires_grad_(True)
    out = m(qg, memg, memg, mask.to(DEV))
    sd = _leafify(_sd(m, "a"))
    qr = q.double().requires_grad_(True)
    memr = qr if kind == "self" else mem.double().requires_grad_(True)
    ref = O.attention(sd, "a", qr, memr, mask, 8, 0.0, False)
    _close(out, ref, 1e-4, "attention out")
    w = torch.randn(B, T, 256)
    (out * w.to(DEV)).sum().backward()
    (ref * w.double()).sum().backward()
    _close(qg.grad, qr.grad, 2e-4, "attention dq")
    if kind == "cross":
        _close(memg.grad, memr.grad, 2e-4, "attention dmem")
    _grads_match(m, sd, "a")


def test_feed_forward_and_combination_forward():
    import gnn_transformer as G
    import fira_oracle as O
    torch.manual_seed(3)
    ff = G.FeedForward(256).to(DEV).eval()
    x = torch.randn(5, 30, 256)
    xg = x.to(DEV).requires_grad_(True)
    out = ff(xg)
    sd = _leafify(_sd(ff, "f"))
    xr = x.double().requires_grad_(True)
    ref = O.feed_forward(sd, "f", xr, 0.0, False)
    _close(out, ref, 1e-4, "ffn out")
    w = torch.randn(5, 30, 256)
    (out * w.to(DEV)).sum().backward(); (ref * w.double()).sum().backward()
    _close(xg.grad, xr.grad, 2e-4, "ffn dx")
    _grads_match(ff, sd, "f")

    comb = G.Combination(8, 256).to(DEV).eval()
    x = torch.randn(3, 210, 256)
    mark_em = torch.randn(3, 210, 256)
    xg, vg = x.to(DEV).requires_grad_(True), mark_em.to(DEV).requires_grad_(True)
    out = comb(xg, xg, vg)                                        # gnn_transformer.py:56
    sd = _leafify(_sd(comb, "c"))
    xr, vr = x.double().requires_grad_(True), mark_em.double().requires_grad_(True)
    ref = O.combination(sd, "c", xr, vr, 8, 0.0, False)
    _close(out, ref, 1e-4, "combination out")
    w = torch.randn(3, 210, 256)
    (out * w.to(DEV)).sum().backward(); (ref * w.double()).sum().backward()
    _close(xg.grad, xr.grad, 2e-4, "combination dx")
    _close(vg.grad, vr.grad, 2e-4, "combination dvalue")
    _grads_match(comb, sd, "c")


def test_combination_layer_forward():
    from combination_layer import CombinationLayer
    g = torch.Generator().manual_seed(4)
    q, k, v = (torch.randn(2, 8, 37, 32, generator=g) for _ in range(3))     # [B, heads, L, d_head], L*heads*B % 8 != 0
    out = CombinationLayer()(q.to(DEV), k.to(DEV), v.to(DEV))
    w = torch.softmax(torch.stack((q * k, q * v), -1).double() / np.sqrt(32), -1)
    ref = w[..., 0] * k.double() + w[..., 1] * v.double()
    _close(out, ref, 1e-5, "combination layer")
