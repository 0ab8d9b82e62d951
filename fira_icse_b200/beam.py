"""Beam search with the reference's exact ranking semantics (run_model.py:187-380), batched on the GPU.

What is kept from the reference:
  * beam 0 starts with probability 1, the others 0; scores are PRODUCTS of probabilities in fp32;
  * a finished beam (last token <eos>) re-enters the ranking with its stored probability through
    `beam_size` extra candidate slots (-1 when absent), unfinished rows of finished samples are -1;
  * candidates = [live beams x (vocab + 210 + 160)] ++ [finished-beam slots], ranked by
    torch.sort(descending=True), top `beam_size` kept; copy ids are mapped back to vocabulary ids
    through the commit's own diff / sub-token ids; the loop stops when every beam of every sample ended.
What changes: the encoder memory is computed once, ALL live beams go through the decoder in one
batched call, and only position `step` is pushed through the output head (the reference recomputes the
full 30 x 25,020 distribution per beam and reads one row of it).  mode="incremental" / "graph" evaluates
only the newest decoder row per step against cached keys/values (incremental.IncrementalDecoder; "graph"
replays each step's kernels as a CUDA graph); mode="full" re-runs the 30-position decoder every step.
The default comes from FIRA_BEAM_MODE (default "graph").  Ranking keeps the reference's candidate layout; only
the first `beam_size` entries of its descending sort are ever used, so the sort is a device top-k.
"""
import os
import weakref

import torch

from .incremental import IncrementalDecoder


_DECODERS = weakref.WeakKeyDictionary()          # model -> {(B, K, ...): IncrementalDecoder}


def _incremental_decoder(model, B, K, tar_len, mem_len, graphs):
    """IncrementalDecoder instances (static buffers, captured graphs) are kept per model and (B, K, mode)."""
    store = _DECODERS.setdefault(model, {})
    key = (B, K, tar_len, mem_len, bool(graphs), model.precision)
    if key not in store:
        store[key] = IncrementalDecoder(model.decoder, B, K, tar_len, mem_len, graphs=graphs)
    return store[key]


@torch.no_grad()
def beam_search(model, sou, mark, ast_change, edge, sub_token, *, beam_size=3, tar_len=30, start_id, eos_id,
                pad_id=0, mode=None):
    """-> (sequences [B, beam, tar_len] int64 padded with pad_id, lengths [B, beam], probs [B, beam])."""
    mode = mode or os.environ.get("FIRA_BEAM_MODE", "graph")
    if mode not in ("full", "incremental", "graph"):
        raise ValueError("beam search mode must be 'full', 'incremental' or 'graph'")
    dev = model.out_fc.weight.device
    sou, mark, ast_change, sub_token = (t.to(dev) for t in (sou, mark, ast_change, sub_token))
    B, K = sou.shape[0], beam_size
    V, n_code = model.vocab_size, sou.shape[1]
    C = V + n_code + sub_token.shape[1]
    memory = model.encoder.encode_memory(sou, mark, ast_change, edge, sub_token)        # once per batch
    mem_mask = torch.cat((sou != pad_id, sub_token != 0), dim=1)
    copy_src = torch.cat((sou, sub_token), dim=1)                                       # copy id -> vocabulary id

    seq = torch.full((B, K, tar_len), pad_id, dtype=torch.long, device=dev)
    seq[:, :, 0] = start_id
    length = torch.ones((B, K), dtype=torch.long, device=dev)
    prob = torch.zeros((B, K), dtype=torch.float32, device=dev)
    prob[:, 0] = 1.0
    ar = torch.arange(B, device=dev)
    inc = None
    if mode != "full":
        inc = _incremental_decoder(model, B, K, tar_len, memory.shape[1], mode == "graph").start(memory, mem_mask)

    for step in range(tar_len - 1):
        last = seq.gather(2, (length - 1).unsqueeze(-1)).squeeze(-1)
        finished = last == eos_id                                                       # [B, K]
        live = [j for j, done in enumerate(finished.all(0).tolist()) if not done]      # one host sync per step
        if not live:
            break
        n_live = len(live)
        live_t = torch.tensor(live, device=dev)
        mem_rep = memory.unsqueeze(1).expand(B, n_live, -1, -1).reshape(B * n_live, memory.shape[1], -1)
        mask_rep = mem_mask.unsqueeze(1).expand(B, n_live, -1).reshape(B * n_live, -1)
        if inc is None:
            tokens = seq[:, live_t].reshape(B * n_live, tar_len)
            dec = model.decoder(tokens, mem_rep, mask_rep, tokens != pad_id)[:, step:step + 1]   # only row `step`
        else:                                          # newest row of every beam against the K/V caches
            row = inc.step(seq[:, :, step].reshape(B * K), step, pad_id)
            dec = row.view(B, K, -1)[:, live_t].reshape(B * n_live, 1, -1)
        gen = torch.softmax(model.out_fc(dec), dim=-1)
        copy, gate = model.copy_net(mem_rep, dec)
        copy = torch.softmax(copy.masked_fill(~mask_rep.unsqueeze(1), -1e9), dim=-1)
        dist = torch.cat((gate[:, :, 0:1] * gen, gate[:, :, 1:2] * copy), dim=-1).view(B, n_live, C)
        dist = dist * prob[:, live_t].unsqueeze(-1)
        dist = dist.masked_fill(finished[:, live_t].unsqueeze(-1), -1.0)
        # finished beams, in beam order, padded with -1 (run_model.py:284-298)
        order = torch.argsort((~finished).to(torch.int8), dim=1, stable=True)            # finished first, stable
        n_fin = finished.sum(1, keepdim=True)
        slot_ok = torch.arange(K, device=dev).unsqueeze(0) < n_fin
        ends_prob = torch.where(slot_ok, prob.gather(1, order), torch.full_like(prob, -1.0))
        cand = torch.cat((dist.view(B, n_live * C), ends_prob), dim=1)
        top_p, top_i = torch.topk(cand, K, dim=-1)          # == sort(descending=True)[:K] (run_model.py:300-303)
        which_beam = top_i // C
        which_tok = top_i % C
        carried = which_beam == n_live                                                   # "keep a finished beam"
        src_beam = torch.where(carried, order.gather(1, which_tok.clamp(max=K - 1)),
                               live_t[which_beam.clamp(max=n_live - 1)])
        tok = torch.where(which_tok >= V, copy_src.gather(1, (which_tok - V).clamp(min=0, max=copy_src.shape[1] - 1)),
                          which_tok)
        new_seq = seq[ar.unsqueeze(1), src_beam]                                         # [B, K, T]
        new_len = length.gather(1, src_beam)
        grow = ~carried
        pos = new_len.clamp(max=tar_len - 1)
        cur = new_seq.gather(2, pos.unsqueeze(-1)).squeeze(-1)
        new_seq.scatter_(2, pos.unsqueeze(-1), torch.where(grow, tok, cur).unsqueeze(-1))
        seq, length, prob = new_seq, new_len + grow.long(), top_p
        if inc is not None:
            inc.reorder((ar.unsqueeze(1) * K + src_beam).reshape(-1))
    return seq, length, prob


def best_sequences(seq, length, prob):
    """run_model.py:351: the beam with the largest probability (first one on ties, like np.argmax)."""
    best = torch.argmax(prob, dim=1)
    ar = torch.arange(seq.shape[0], device=seq.device)
    return seq[ar, best], length[ar, best]
