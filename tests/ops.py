"""Operator-level Python wrappers over the C ABI (torch tensors in, torch tensors out).

TEST / BENCH infrastructure (tests/test_ops_gpu.py, tools/bench_ops.py): the model path (wenet_b200/asr_model.py) calls
the stage-level entry points, never these.  Every function requires CUDA tensors — there is no fallback.
"""
import math

import torch

from wenet_b200 import _lib
from wenet_b200._lib import check, cur_stream, ptr

EPI_BF16, EPI_BF16_SILU, EPI_BF16_RELU, EPI_RESID_F32, EPI_GLU_BF16, EPI_F32 = range(6)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.WbError("wenet_b200 ops need CUDA tensors (no CPU fallback)")


def gemm(a, b, bias=None, epi=EPI_BF16, alpha=1.0, out=None, split3=False):
    """out = epi(a @ b.T + bias); a [M,K] bf16, b [N,K] bf16."""
    _need_cuda(a, b, bias, out)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    M, K = a.shape
    N = b.shape[0]
    assert b.shape[1] == K and a.stride(1) == 1 and b.is_contiguous()
    on = N // 2 if epi == EPI_GLU_BF16 else N
    if out is None:
        if epi in (EPI_RESID_F32, EPI_F32):
            out = torch.zeros(M, on, device=a.device, dtype=torch.float32)
        else:
            out = torch.empty(M, on * (3 if split3 else 1), device=a.device, dtype=torch.bfloat16)
    check(_lib.load().wb_op_gemm(ptr(a), a.stride(0), ptr(b), M, N, K, ptr(bias), epi, float(alpha),
                                 ptr(out), out.stride(0), int(split3), cur_stream()), "wb_op_gemm")
    return out


def gemm_resid_ln(a, b, bias, x, gamma, beta, alpha=1.0, eps=1e-5, gamma1=None, beta1=None):
    """x += alpha * (a @ b.T + bias) in place; returns bf16 LayerNorm(x).  With gamma1 / beta1: x = LayerNorm_1(x + ...)
    in place and the result is LayerNorm(gamma, beta)(x)."""
    _need_cuda(a, b, bias, x, gamma, beta, gamma1, beta1)
    M, K = a.shape
    N = b.shape[0]
    out = torch.empty(M, N, device=a.device, dtype=torch.bfloat16)
    check(_lib.load().wb_op_gemm_resid_ln(ptr(a), a.stride(0), ptr(b), M, N, K, ptr(bias), float(alpha), ptr(x), x.stride(0),
                                          ptr(gamma1), ptr(beta1), ptr(gamma), ptr(beta), float(eps), ptr(out), out.stride(0), cur_stream()),
          "wb_op_gemm_resid_ln")
    return out


def layernorm(x, gamma, beta, eps=1e-5, want_bf16=True, want_f32=False, split3=False):
    _need_cuda(x, gamma, beta)
    M, d = x.shape
    ob = torch.empty(M, d * (3 if split3 else 1), device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    of = torch.empty(M, d, device=x.device, dtype=torch.float32) if want_f32 else None
    check(_lib.load().wb_op_layernorm(ptr(x), x.stride(0), M, d, ptr(gamma), ptr(beta), float(eps),
                                      ptr(ob), ob.stride(0) if ob is not None else 0, int(split3),
                                      ptr(of), of.stride(0) if of is not None else 0, cur_stream()),
          "wb_op_layernorm")
    return ob, of


def attention(q, k, v, q_start, q_len, k_start, k_len, heads, kbias=None, chunk_size=0,
              num_left_chunks=-1, scale=None, q_col0=0, k_col0=0, v_col0=0, v_mode=0, max_q_len=None):
    """q/k/v: 2-D bf16 row-major buffers, head h of q at columns [q_col0 + 64 h, +64)."""
    _need_cuda(q, k, v, kbias)
    if scale is None:
        scale = 1.0 / math.sqrt(64.0)
    batch = q_start.numel()
    if max_q_len is None:
        max_q_len = int(q_len.max().item())
    out = torch.zeros(q.shape[0], heads * 64, device=q.device, dtype=torch.bfloat16)
    check(_lib.load().wb_op_attention(
        ptr(q), q.stride(0), q.shape[0], q_col0, ptr(k), k.stride(0), k.shape[0], k_col0,
        ptr(v), v.stride(0), v.shape[0], v_col0, ptr(kbias), kbias.stride(0) if kbias is not None else 0,
        ptr(q_start), ptr(q_len), ptr(k_start), ptr(k_len), batch, heads, max_q_len, chunk_size,
        num_left_chunks, float(scale), ptr(out), out.stride(0), 0, v_mode, cur_stream()), "wb_op_attention")
    return out


def relpos_kprep(k, pos_proj, row_pos, bias_u, bias_v, heads):
    _need_cuda(k, pos_proj, row_pos)
    M = k.shape[0]
    kp = torch.empty(M, heads * 64, device=k.device, dtype=torch.bfloat16)
    kb = torch.empty(M, heads, device=k.device, dtype=torch.float32)
    check(_lib.load().wb_op_relpos_kprep(ptr(k), k.stride(0), ptr(pos_proj), ptr(row_pos), ptr(bias_u),
                                         ptr(bias_v), M, heads, ptr(kp), kp.stride(0), ptr(kb),
                                         cur_stream()), "wb_op_relpos_kprep")
    return kp, kb


def dwconv(g, seq_start, seq_len, out_start, w, bias, gamma, beta, ksize, causal, norm_type=0, eps=1e-5,
           lead=0, pad_vec=None, pad_until=0, out_rows=None):
    _need_cuda(g, w)
    d = g.shape[1]
    batch = seq_start.numel()
    max_len = int(seq_len.max().item())
    if out_rows is None:
        out_rows = g.shape[0]
    out = torch.zeros(out_rows, d, device=g.device, dtype=torch.bfloat16)
    check(_lib.load().wb_op_dwconv(ptr(g), g.stride(0), ptr(seq_start), ptr(seq_len), ptr(out_start), batch,
                                   max_len, lead, d, ksize, int(causal), ptr(w), ptr(bias), norm_type,
                                   ptr(gamma), ptr(beta), float(eps), ptr(pad_vec), pad_until, ptr(out),
                                   out.stride(0), cur_stream()), "wb_op_dwconv")
    return out


def logsoftmax_topk(logits, V, topk, blank_id=0, blank_penalty=0.0):
    """in-place log-softmax over the first V columns of logits [M, ld]; returns (topk_val, topk_idx)."""
    _need_cuda(logits)
    M = logits.shape[0]
    tv = torch.empty(M, max(topk, 1), device=logits.device, dtype=torch.float32)
    ti = torch.empty(M, max(topk, 1), device=logits.device, dtype=torch.int32)
    check(_lib.load().wb_op_logsoftmax_topk(ptr(logits), logits.stride(0), M, V, blank_id, float(blank_penalty),
                                            topk, ptr(tv), ptr(ti), cur_stream()), "wb_op_logsoftmax_topk")
    return tv, ti


def lse_topk(logits, V, topk, blank_id=0, blank_penalty=0.0):
    """(topk_val, topk_idx) of log_softmax(logits[:, :V]) without touching `logits`."""
    _need_cuda(logits)
    M = logits.shape[0]
    tv = torch.empty(M, topk, device=logits.device, dtype=torch.float32)
    ti = torch.empty(M, topk, device=logits.device, dtype=torch.int32)
    check(_lib.load().wb_op_lse_topk(ptr(logits), logits.stride(0), M, V, blank_id, float(blank_penalty), topk, ptr(tv),
                                     ptr(ti), cur_stream()), "wb_op_lse_topk")
    return tv, ti


def ctc_greedy_search(topk_idx, seq_start, seq_len, blank_id=0):
    _need_cuda(topk_idx)
    batch = seq_start.numel()
    max_len = int(seq_len.max().item()) if batch else 0
    toks = torch.zeros(batch, max(max_len, 1), device=topk_idx.device, dtype=torch.int32)
    lens = torch.zeros(batch, device=topk_idx.device, dtype=torch.int32)
    check(_lib.load().wb_ctc_greedy_search(ptr(topk_idx), topk_idx.stride(0), ptr(seq_start), ptr(seq_len), batch,
                                           blank_id, ptr(toks), toks.stride(0), ptr(lens), cur_stream()),
          "wb_ctc_greedy_search")
    return toks, lens


def ctc_prefix_beam_search(topk_val, topk_idx, seq_start, seq_len, beam, blank_id=0, max_len=None):
    _need_cuda(topk_val, topk_idx)
    batch = seq_start.numel()
    if max_len is None:
        max_len = max(int(seq_len.max().item()), 1)
    dev = topk_val.device
    toks = torch.zeros(batch, beam, max_len, device=dev, dtype=torch.int32)
    times = torch.zeros(batch, beam, max_len, device=dev, dtype=torch.int32)
    lens = torch.zeros(batch, beam, device=dev, dtype=torch.int32)
    scores = torch.zeros(batch, beam, device=dev, dtype=torch.float64)
    nhyp = torch.zeros(batch, device=dev, dtype=torch.int32)
    lib = _lib.load()
    wsb = lib.wb_prefix_beam_workspace_bytes(batch, beam, max_len)
    ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
    check(lib.wb_ctc_prefix_beam_search(ptr(topk_val), ptr(topk_idx), topk_val.stride(0), ptr(seq_start),
                                        ptr(seq_len), batch, beam, blank_id, max_len, ptr(toks), ptr(times),
                                        ptr(lens), ptr(scores), ptr(nhyp), ptr(ws), wsb, cur_stream()),
          "wb_ctc_prefix_beam_search")
    return toks, times, lens, scores, nhyp
