// CTC prefix beam search on the GPU: ONE CTA (9 warps) PER UTTERANCE, frames sequential, all (token, prefix) pairs
// and all candidates of a frame in parallel.  Bit-level restatement of the reference's host loop
//   wenet/models/transformer/search.py:127-249 (ctc_prefix_beam_search, PrefixScore :64-106)
//   wenet/utils/common.py:302-310 (log_add, Python float == IEEE double)
// including: double-precision scores, Viterbi scores / token times (times_s / times_ns / cur_token_prob rules of
// :166-219), dict insertion order as the tie-break of the stable `sorted(..., reverse=True)` (:222-225),
// first-max-wins comparisons.
//
// Data structures
//   * prefixes are nodes of a CANONICAL trie kept in a per-utterance pool in HBM: every distinct prefix string has
//     exactly one node (a new beam entry first looks its (parent node, token) up in the parent's child list and only
//     creates a node when none exists), so "prefix + u is already in the beam" is the exact integer test
//     parent_node[q] == node[p] && last[q] == u.  (Round 1 used a 64-bit rolling hash of the string as identity.)
//   * times lists are persistent (immutable) linked lists (prev, frame) in a second pool, so `list.copy()` / `append`
//     / `[-1] = t` of the reference are O(1) pointer operations
//   * the beam (<= 16 entries) and the <= beam + beam^2 candidates of a frame live in shared memory.
// Thread mapping per frame (288 threads = 9 warps; thread c owns candidate slot c):
//   P1  thread (ui, pi) = (tid / 16, tid % 16): extension candidate of prefix pi by top-k token ui
//   P2  thread q < beam size: the "unchanged prefix" candidate q; its (<= 3) updates (blank, repeat, extension of its
//       parent landing on it) are replayed in the reference's loop order (token-major, prefix-minor), so every
//       floating-point operation happens in the same order as on the host
//   P3  warp 0 ranks its own candidates (unchanged prefixes + extensions by the best token) by shuffle rotations on an
//       order-preserving integer image of the fp64 total score and publishes its beam-th best as a pruning bound;
//       thread c = candidate slot of the other warps only compares with the bound; the (typically 10-15) survivors
//       are ranked by warp 0 the same way -> exact top-beam in the reference's stable-sort order.  P2 and the key
//       set-up of the other warps overlap.
//   P4  thread r < new beam size: materialise entry r (canonical trie node, times lists)
#include "common.cuh"
#include "kernels.h"
#include <limits.h>
#include <math_constants.h>
#include "softplus_table.h"

namespace wb {

namespace {

constexpr int MAXB = 16;
constexpr int NCAND = MAXB + MAXB * MAXB;

__device__ __forceinline__ double neg_inf() { return -CUDART_INF; }

// f(d) = log(1 + exp(d)) for d <= 0: piecewise degree-12 polynomials (softplus_table.h, max abs error 1.1e-16
// against the exact value — the same as libm's log1p(exp(d))).  libdevice's fp64 exp + log are ~250 dependent
// instructions; in this single-warp, latency-bound kernel they were 2/3 of the run time (ncu, profiles/).
__device__ __forceinline__ double softplus_neg(double d) {
    // reference arithmetic: 1.0 + exp(d) rounds to 1.0 once exp(d) <= 2^-53, and log(1.0) == 0 exactly
    if (d < -36.7368005696771) return 0.0;
    int idx = (int)(-2.0 * d);
    idx = idx > WB_SOFTPLUS_NINT - 1 ? WB_SOFTPLUS_NINT - 1 : idx;
    const double t = fma(4.0, d, 2.0 * (double)idx + 1.0);   // (d - mid) / 0.25, mid = -(idx + 0.5) / 2
    const double* c = g_softplus_tab[idx];
    double k[WB_SOFTPLUS_DEG + 1];
#pragma unroll
    for (int j = 0; j <= WB_SOFTPLUS_DEG; ++j) k[j] = __ldg(c + j);
    // Estrin's scheme (5 dependent fma levels instead of 12: this chain sits on the per-frame critical path of the beam
    // search); the polynomial is accurate to 1.1e-16, the evaluation order changes the result by at most an ulp
    static_assert(WB_SOFTPLUS_DEG == 12, "Estrin scheme below is written for degree 12");
    const double t2 = t * t, t4 = t2 * t2, t8 = t4 * t4;
    const double p01 = fma(k[1], t, k[0]), p23 = fma(k[3], t, k[2]), p45 = fma(k[5], t, k[4]), p67 = fma(k[7], t, k[6]);
    const double p89 = fma(k[9], t, k[8]), pab = fma(k[11], t, k[10]);
    const double q0 = fma(p23, t2, p01), q1 = fma(p67, t2, p45), q2 = fma(pab, t2, p89);
    const double r0 = fma(q1, t4, q0), r1 = fma(k[12], t4, q2);
    return fma(r1, t8, r0);
}

// log_add of wenet/utils/common.py:302-310 for two arguments:
//   m + log(exp(a - m) + exp(b - m)) = m + log(1 + exp(lo - m))   (the larger argument contributes exp(0) == 1.0
//   exactly); log_add(-inf, x) == x exactly, so those cases return early as the reference's arithmetic does.
__device__ __forceinline__ double log_add2(double a, double b) {
    if (a == neg_inf()) return b;
    if (b == neg_inf()) return a;
    const double m = a > b ? a : b;
    const double lo = a > b ? b : a;
    return m + softplus_neg(lo - m);
}

struct Beam {
    double s[MAXB], ns[MAXB], vs[MAXB], vns[MAXB];
    double score[MAXB], vit[MAXB];
    int len[MAXB], last[MAXB], node[MAXB], par[MAXB], ts[MAXB], tns[MAXB], times[MAXB];
    double ctx_score[MAXB];   // context-graph biasing (search.py:76-78): accumulated bonus and graph state
    int ctx_state[MAXB];
    int n;
};

struct Cand {
    double s[NCAND], ns[NCAND], vs[NCAND], vns[NCAND], total[NCAND];
    int len[NCAND], last[NCAND];
    int node[NCAND];       // existing trie node (unchanged prefix) or PARENT node (extension)
    int par[NCAND];        // parent node of an unchanged prefix
    int new_tok[NCAND];    // >= 0: extension by this token (needs a trie node)
    int ts[NCAND];         // times_s head
    int tns[NCAND];        // times_ns head if tns_new == 0, else prev pointer of the node to create
    int tns_new[NCAND];    // 1: times_ns = list(tns) + [t]
    int first[NCAND];      // first-touch sequence number == dict insertion order
    int valid[NCAND];      // frame stamp (t + 1)
    double ctx_score[NCAND];
    int ctx_state[NCAND];
};

struct PbDev {
    const float* topk_val;
    const int* topk_idx;
    int topk;
    const int* seq_start;
    const int* seq_len;
    int beam, blank_id, max_len;
    int* out_tokens;
    int* out_times;
    int* out_lens;
    double* out_scores;
    int* out_nhyp;
    int* pool;  // per utterance: [6][max_len * beam] ints: trie parent, token, first child, next sibling; time prev, frame
    // context graph (num_nodes == 0: no biasing): wenet/utils/context_graph.py flattened, node 0 = root
    int cg_nodes;
    const int* cg_child_off;
    const int* cg_child_tok;
    const int* cg_child_node;
    const int* cg_fail;
    const int* cg_token;
    const double* cg_node_score;
    const double* cg_token_score;
    const double* cg_output_score;
};

// child of `node` labelled `token` (children are sorted by token), -1 if none
__device__ __forceinline__ int cg_child(const PbDev& P, int node, int token) {
    int lo = __ldg(P.cg_child_off + node), hi = __ldg(P.cg_child_off + node + 1);
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const int tk = __ldg(P.cg_child_tok + mid);
        if (tk == token) return __ldg(P.cg_child_node + mid);
        if (tk < token) lo = mid + 1;
        else hi = mid;
    }
    return -1;
}
// ContextGraph.forward_one_step (context_graph.py:212-247): score of moving from `state` by `token`, and the new state
__device__ __forceinline__ double cg_step(const PbDev& P, int state, int token, int* next_state) {
    int node = cg_child(P, state, token);
    double score;
    if (node >= 0) {
        score = __ldg(P.cg_token_score + node);
    } else {
        node = __ldg(P.cg_fail + state);
        while (cg_child(P, node, token) < 0) {
            node = __ldg(P.cg_fail + node);
            if (__ldg(P.cg_token + node) == -1) break;   // root
        }
        const int c2 = cg_child(P, node, token);
        if (c2 >= 0) node = c2;
        score = __ldg(P.cg_node_score + node) - __ldg(P.cg_node_score + state);   // the score of the fail path
    }
    *next_state = node;
    return score + __ldg(P.cg_output_score + node);
}

// order-preserving image of a double in the unsigned integers (-inf < ... < -0 == +0 < ... ; no NaNs occur)
__device__ __forceinline__ unsigned long long order_key(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v + 0.0);   // -0.0 + 0.0 == +0.0
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
// candidate a ranks before b: total desc, then dict insertion order asc
__device__ __forceinline__ bool key_better(unsigned long long ka, int fa, unsigned long long kb, int fb) {
    return ka > kb || (ka == kb && fa < fb);
}
// rank of this lane's (key, first) among lanes [0, n) of the warp (n warp-uniform; lanes >= n get a meaningless value).
// Real candidates have pairwise different `first`; empty lanes carry key 0 / first INT_MAX and are never better.
// n - 1 independent shuffle rotations (not a sorting network: no dependent stages).
__device__ __forceinline__ int warp_rank(unsigned long long key, int first, int lane, int n) {
    const unsigned klo = (unsigned)key, khi = (unsigned)(key >> 32);
    int rk = 0;
#pragma unroll 4
    for (int r = 1; r < n; ++r) {
        int src = lane + r;
        src = src >= n ? src - n : src;
        const unsigned olo = __shfl_sync(0xffffffffu, klo, src);
        const unsigned ohi = __shfl_sync(0xffffffffu, khi, src);
        const int of = __shfl_sync(0xffffffffu, first, src);
        const unsigned long long ok = ((unsigned long long)ohi << 32) | olo;
        rk += key_better(ok, of, key, first) ? 1 : 0;
    }
    return rk;
}

constexpr int PB_SLOT_WARPS = (NCAND + 31) / 32;   // warps that own candidate slots in P3 (9: 272 slots at beam 16)
constexpr int PB_THREADS = 32 * PB_SLOT_WARPS;     // 288: thread c owns candidate slot c; threads < 256 are the (ui, pi) pairs

__global__ void __launch_bounds__(PB_THREADS, 3)
prefix_beam_kernel(PbDev P) {
    __shared__ Beam Bs[2];
    __shared__ Cand C;
    __shared__ float tk_val[2][MAXB];
    __shared__ int tk_idx[2][MAXB];
    __shared__ int dest[MAXB * MAXB];
    __shared__ int rank_slot[MAXB];
    __shared__ unsigned long long thr_key[1];
    __shared__ int thr_first[1];
    __shared__ int warp_valid[PB_SLOT_WARPS];
    __shared__ unsigned long long sv_key[NCAND];
    __shared__ int sv_first[NCAND], sv_slot[NCAND];
    __shared__ int n_surv;
    __shared__ int root_child;      // head of the root's child list (the root has no pool entry)
    // P1 -> P2 hand-over, frame-stamped ((t + 1) * 32 + index) so that nothing has to be cleared per frame:
    __shared__ int s_ui_blank;      // top-k position of <blank>
    __shared__ int s_ui_last[MAXB]; // top-k position of prefix q's last token
    __shared__ int s_pe[MAXB];      // the prefix whose extension lands on beam entry q

    const int utt = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int beam = P.beam;
    const int T = P.seq_len[utt];
    const long long f0 = P.seq_start[utt];
    const long long pool_n = (long long)P.max_len * beam;
    int* trie_parent = P.pool + (long long)utt * 6 * pool_n;
    int* trie_tok = trie_parent + pool_n;
    int* trie_child = trie_tok + pool_n;     // first child
    int* trie_sib = trie_child + pool_n;     // next sibling
    int* time_prev = trie_sib + pool_n;
    int* time_t = time_prev + pool_n;
    const int ncs = MAXB + beam * MAXB;      // candidate slots in use

    for (int c = tid; c < NCAND; c += PB_THREADS) C.valid[c] = 0;
    if (tid == 0) {
        Beam& B = Bs[0];
        B.n = 1;
        B.s[0] = 0.0;
        B.ns[0] = neg_inf();
        B.vs[0] = 0.0;
        B.vns[0] = 0.0;
        B.score[0] = 0.0;   // log_add(0, -inf)
        B.vit[0] = 0.0;     // v_s > v_ns is false -> v_ns
        B.times[0] = -1;
        B.len[0] = 0;
        B.last[0] = -1;
        B.node[0] = -1;     // the root
        B.par[0] = -2;      // the root is nobody's child
        B.ts[0] = -1;
        B.tns[0] = -1;
        B.ctx_score[0] = 0.0;
        B.ctx_state[0] = 0;
        root_child = -1;
        n_surv = 0;
        s_ui_blank = 0;
    }
    if (tid < MAXB) {
        s_ui_last[tid] = 0;
        s_pe[tid] = 0;
    }
    if (tid < beam && T > 0) {
        tk_val[0][tid] = P.topk_val[f0 * P.topk + tid];
        tk_idx[0][tid] = P.topk_idx[f0 * P.topk + tid];
    }
    __syncthreads();

    int cur = 0;
    for (int t = 0; t < T; ++t) {
        const float* tkv = tk_val[t & 1];
        const int* tki = tk_idx[t & 1];
        if (tid < beam && t + 1 < T) {   // next frame's top-k lands in the other buffer behind this frame's work
            tk_val[(t + 1) & 1][tid] = P.topk_val[(f0 + t + 1) * P.topk + tid];
            tk_idx[(t + 1) & 1][tid] = P.topk_idx[(f0 + t + 1) * P.topk + tid];
        }
        Beam& B = Bs[cur];
        const int nb = B.n;

        // ---- P1: extension candidates, one thread per (token ui, prefix pi) ----
        {
            const int ui = tid >> 4, pi = tid & (MAXB - 1);
            if (tid < MAXB * MAXB && ui < beam && pi < nb) {
                const int u = tki[ui];
                const int stamp = (t + 1) * 32;
                int d = -1;
                if (u == P.blank_id) {
                    if (pi == 0) s_ui_blank = stamp + ui;
                } else {
                    const int np = B.node[pi];
                    if (u == B.last[pi]) s_ui_last[pi] = stamp + ui;
                    d = MAXB + ui * MAXB + pi;
#pragma unroll
                    for (int q = 0; q < MAXB; ++q)   // (fixed trip count: all loads issue up front)
                        if (q < nb && B.par[q] == np && B.last[q] == u) d = q;   // prefix pi + u is beam entry q (canonical trie)
                    if (d < MAXB) s_pe[d] = stamp + pi;    // at most one (prefix, token) pair lands on a given entry
                    if (d >= MAXB) {
                        const double prob = (double)tkv[ui];
                        const bool rep = (u == B.last[pi]);
                        C.s[d] = neg_inf();
                        C.vs[d] = neg_inf();
                        C.ts[d] = -1;
                        // log_add(-inf, x) == x exactly
                        const double nsv = (rep ? B.s[pi] : B.score[pi]) + prob;
                        double vn = (rep ? B.vs[pi] : B.vit[pi]) + prob;
                        // reference: `if next.v_ns < y` with next.v_ns = -inf: false only when y == -inf
                        if (vn > neg_inf()) {
                            C.tns[d] = rep ? B.ts[pi] : B.times[pi];
                            C.tns_new[d] = 1;
                        } else {
                            vn = neg_inf();
                            C.tns[d] = -1;
                            C.tns_new[d] = 0;
                        }
                        C.ns[d] = nsv;
                        C.vns[d] = vn;
                        C.len[d] = B.len[pi] + 1;
                        C.last[d] = u;
                        C.node[d] = np;
                        C.new_tok[d] = u;
                        C.first[d] = (ui * nb + pi) * 2 + (rep ? 1 : 0);
                        C.total[d] = nsv;  // log_add(-inf, ns)
                        if (P.cg_nodes > 0) {   // update_context (search.py:101-106): the only touch of a new prefix
                            int nst;
                            const double sc = cg_step(P, B.ctx_state[pi], u, &nst);
                            C.ctx_score[d] = B.ctx_score[pi] + sc;
                            C.ctx_state[d] = nst;
                        }
                        C.valid[d] = t + 1;   // frame stamp: no per-frame reset of the flags
                    }
                }
                dest[ui * MAXB + pi] = d;
            }
        }
        __syncthreads();

        // ---- P2 (warp 0 only) + P3: second beam prune = top-beam by (total desc, dict insertion order asc) ----
        // Warp 0 owns the unchanged prefixes (slots 0..15) and the extensions by the best token (slots 16..31): it runs
        // P2, ranks its own <= 2 nb candidates and publishes its beam-th best as THE pruning bound (a lower bound of the
        // global beam-th best; with blank-dominated CTC posteriors it is nearly tight).  The other warps meanwhile
        // load and key their slots; after the barrier they only compare against the bound.
        unsigned long long my_key = 0ull;
        int my_first = INT_MAX;
        bool my_valid = false;
        if (warp == 0) {
            if (tid < nb) {
                const int q = tid;
                const int lastq = B.last[q];
                const int stamp = (t + 1) * 32;
                const int vb = s_ui_blank, vl = s_ui_last[q], vp = s_pe[q];
                const int ui_blank = (vb >= stamp) ? vb - stamp : -1;
                const int ui_last = (vl >= stamp) ? vl - stamp : -1;
                double s = neg_inf(), ns = neg_inf(), vs = neg_inf(), vns = neg_inf();
                int ts = -1, tns = -1, tns_new = 0, first = INT_MAX, any = 0;
                if (ui_blank >= 0) {
                    const double prob = (double)tkv[ui_blank];
                    s = B.score[q] + prob;          // log_add(-inf, x)
                    vs = B.vit[q] + prob;
                    ts = B.times[q];
                    first = (ui_blank * nb + q) * 2;
                    any = 1;
                }
                if (ui_last >= 0) {
                    const double prob = (double)tkv[ui_last];
                    const int pe = (vp >= stamp) ? vp - stamp : -1;  // parent prefix whose extension by last(q) lands on q
                    bool cur_set = false;
                    // two possible events on ns, in prefix order: extension from pe, repeat from q itself
                    for (int ev = 0; ev < 2; ++ev) {
                        const bool do_ext = (pe >= 0) && ((ev == 0) == (pe < q));
                        const bool do_rep = (ev == 0) == !(pe >= 0 && pe < q);
                        if (do_ext) {
                            const bool rep = (lastq == B.last[pe]);
                            ns = log_add2(ns, (rep ? B.s[pe] : B.score[pe]) + prob);
                            const double y = (rep ? B.vs[pe] : B.vit[pe]) + prob;
                            if (vns < y) {
                                vns = y;
                                cur_set = true;
                                tns = rep ? B.ts[pe] : B.times[pe];
                                tns_new = 1;
                            }
                            first = min(first, (ui_last * nb + pe) * 2 + (rep ? 1 : 0));
                        } else if (do_rep) {
                            ns = log_add2(ns, B.ns[q] + prob);
                            const double y = B.vns[q] + prob;
                            if (vns < y) {
                                vns = y;
                                if (!cur_set) {
                                    cur_set = true;
                                    // times_ns = prefix.times_ns.copy(); times_ns[-1] = t
                                    const int hd = B.tns[q];
                                    tns = (hd >= 0) ? time_prev[hd] : -1;
                                    tns_new = 1;
                                }
                            }
                            first = min(first, (ui_last * nb + q) * 2);
                        }
                    }
                    any = 1;
                }
                if (any) {
                    C.s[q] = s;
                    C.ns[q] = ns;
                    C.vs[q] = vs;
                    C.vns[q] = vns;
                    C.ts[q] = ts;
                    C.tns[q] = tns;
                    C.tns_new[q] = tns_new;
                    C.len[q] = B.len[q];
                    C.last[q] = lastq;
                    C.node[q] = B.node[q];
                    C.par[q] = B.par[q];
                    C.new_tok[q] = -1;
                    C.first[q] = first;
                    C.total[q] = log_add2(s, ns);
                    if (P.cg_nodes > 0) {
                        // `has_context`: the FIRST touch of next_hyps[prefix] fixes its context (search.py:171-173,
                        // 200-203).  Touches by q itself (blank, repeat) copy q's context; the extension of its parent
                        // pe by last(q) runs update_context(pe) - it is first iff its sequence number is the minimum.
                        const int pe = (vp >= stamp) ? vp - stamp : -1;
                        const bool rep_pe = pe >= 0 && lastq == B.last[pe];
                        const int f_ext = pe >= 0 ? (ui_last * nb + pe) * 2 + (rep_pe ? 1 : 0) : INT_MAX;
                        if (pe >= 0 && ui_last >= 0 && f_ext == first) {
                            int nst;
                            const double sc = cg_step(P, B.ctx_state[pe], lastq, &nst);
                            C.ctx_score[q] = B.ctx_score[pe] + sc;
                            C.ctx_state[q] = nst;
                        } else {
                            C.ctx_score[q] = B.ctx_score[q];
                            C.ctx_state[q] = B.ctx_state[q];
                        }
                    }
                    C.valid[q] = t + 1;
                }
            }
            __syncwarp();
            // compact view of warp 0's candidates: lane l < nb -> slot l, nb <= l < 2 nb -> slot 16 + (l - nb)
            const int n0 = 2 * nb;
            const int c = lane < nb ? lane : (MAXB + lane - nb);
            my_valid = lane < n0 && C.valid[c] == t + 1;
            if (my_valid) {
                my_key = order_key(P.cg_nodes > 0 ? C.total[c] + C.ctx_score[c] : C.total[c]);   // total_score()
                my_first = C.first[c];
            }
            const int rk = warp_rank(my_key, my_first, lane, n0);
            const unsigned vm = __ballot_sync(0xffffffffu, my_valid);
            const unsigned bm = __ballot_sync(0xffffffffu, my_valid && rk == beam - 1);
            if (lane == 0) {
                warp_valid[0] = __popc(vm);
                if (!bm) thr_key[0] = 0ull, thr_first[0] = INT_MAX;   // fewer than beam candidates in warp 0: no bound
            }
            if (my_valid && rk == beam - 1) thr_key[0] = my_key, thr_first[0] = my_first;
            if (my_valid && rk < beam) {   // warp 0's survivors
                const int pos = atomicAdd(&n_surv, 1);
                sv_key[pos] = my_key;
                sv_first[pos] = my_first;
                sv_slot[pos] = c;
            }
        } else {
            const int c = tid;     // slots 32 .. ncs - 1
            my_valid = (c < ncs) && (C.valid[c] == t + 1);
            if (my_valid) {
                my_key = order_key(P.cg_nodes > 0 ? C.total[c] + C.ctx_score[c] : C.total[c]);
                my_first = C.first[c];
            }
            const unsigned vm = __ballot_sync(0xffffffffu, my_valid);
            if (lane == 0) warp_valid[warp] = __popc(vm);
        }
        __syncthreads();
        if (warp != 0 && my_valid && !key_better(thr_key[0], thr_first[0], my_key, my_first)) {
            const int pos = atomicAdd(&n_surv, 1);
            sv_key[pos] = my_key;
            sv_first[pos] = my_first;
            sv_slot[pos] = tid;
        }
        __syncthreads();
        // warp 0 ranks the survivors (any order in the list: the rank is a function of the keys only)
        int nvalid = 0;
#pragma unroll
        for (int w = 0; w < PB_SLOT_WARPS; ++w) nvalid += warp_valid[w];
        const int nnew = nvalid < beam ? nvalid : beam;
        if (warp == 0) {
            const int nsv = n_surv;
            if (nsv <= 32) {
                const unsigned long long k = lane < nsv ? sv_key[lane] : 0ull;
                const int f = lane < nsv ? sv_first[lane] : INT_MAX;
                const int rk = warp_rank(k, f, lane, nsv);
                if (lane < nsv && rk < beam) rank_slot[rk] = sv_slot[lane];
            } else {
                // many candidates tie with / exceed the bound: rank by counting over the list
                for (int i = lane; i < nsv; i += 32) {
                    const unsigned long long k = sv_key[i];
                    const int f = sv_first[i];
                    int rk = 0;
                    for (int j = 0; j < nsv; ++j) rk += key_better(sv_key[j], sv_first[j], k, f) ? 1 : 0;
                    if (rk < beam) rank_slot[rk] = sv_slot[i];
                }
            }
        }
        __syncthreads();

        // ---- P4: materialise the new beam ----
        Beam& NB = Bs[cur ^ 1];
        int new_node = -1, new_parent = -1;   // P4b: node this thread has to link into its parent's child list
        {   // plain field copies, one (entry, field) per thread of warps 1..4 (warp 0 does the pointer work below)
            const int r = tid & (MAXB - 1), fld = (tid >> 4) - 2;
            if (fld >= 0 && fld < 7 && r < nnew) {
                const int c = rank_slot[r];
                switch (fld) {
                    case 0: NB.s[r] = C.s[c]; break;
                    case 1: NB.ns[r] = C.ns[c]; break;
                    case 2: NB.vs[r] = C.vs[c]; break;
                    case 3: NB.vns[r] = C.vns[c]; break;
                    case 4: NB.score[r] = C.total[c]; break;
                    case 5: NB.len[r] = C.len[c]; break;
                    default: NB.last[r] = C.last[c]; break;
                }
            }
        }
        if (tid < nnew) {
            const int r = tid;
            const int c = rank_slot[r];
            const int pool_i = t * beam + r;
            const double vs = C.vs[c], vns = C.vns[c];
            const int tok = C.new_tok[c];
            if (tok >= 0) {
                // canonical node of (parent, tok): reuse the parent's existing child if the string was in a beam before
                const int par = C.node[c];
                int ch = (par >= 0) ? trie_child[par] : root_child;
                while (ch >= 0 && trie_tok[ch] != tok) ch = trie_sib[ch];
                if (ch < 0) {
                    ch = pool_i;
                    trie_parent[ch] = par;
                    trie_tok[ch] = tok;
                    trie_child[ch] = -1;
                    new_node = ch;
                    new_parent = par;
                }
                NB.node[r] = ch;
                NB.par[r] = par;
            } else {
                NB.node[r] = C.node[c];
                NB.par[r] = C.par[c];
            }
            if (P.cg_nodes > 0) {
                NB.ctx_score[r] = C.ctx_score[c];
                NB.ctx_state[r] = C.ctx_state[c];
            }
            const int tsv = C.ts[c];
            int tnsv;
            if (C.tns_new[c]) {
                time_prev[pool_i] = C.tns[c];
                time_t[pool_i] = t;
                tnsv = pool_i;
            } else {
                tnsv = C.tns[c];
            }
            NB.ts[r] = tsv;
            NB.tns[r] = tnsv;
            const bool sb = vs > vns;
            NB.vit[r] = sb ? vs : vns;
            NB.times[r] = sb ? tsv : tnsv;
        }
        if (tid == 0) {
            NB.n = nnew;
            n_surv = 0;
        }
        // P4b: link the new nodes (after every look-up of this frame has finished; siblings of one parent are linked
        // one after the other by the lanes of warp 0 in lane order)
        if (warp == 0) {
            __syncwarp();
            const unsigned nm = __ballot_sync(0xffffffffu, new_node >= 0);
            for (unsigned m = nm; m; m &= m - 1) {
                const int src = __ffs(m) - 1;
                if (lane == src) {
                    if (new_parent >= 0) {
                        trie_sib[new_node] = trie_child[new_parent];
                        trie_child[new_parent] = new_node;
                    } else {
                        trie_sib[new_node] = root_child;
                        root_child = new_node;
                    }
                }
                __syncwarp();
            }
        }
        __syncthreads();
        cur ^= 1;
    }

    // ---- emit n-best ----
    const Beam& B = Bs[cur];
    const int nb = B.n;
    if (tid == 0) P.out_nhyp[utt] = nb;
    if (tid < nb) {
        const int r = tid;
        const long long o = ((long long)utt * beam + r) * P.max_len;
        const int ln = B.len[r];
        P.out_lens[utt * beam + r] = ln;
        // search.py:229-248: finalize() REPLACES the accumulated context score by the implicit fail arc to the root
        // (-node_score of the final state); without a graph total_score() == score()
        P.out_scores[utt * beam + r] =
            P.cg_nodes > 0 ? B.score[r] + (-__ldg(P.cg_node_score + B.ctx_state[r])) : B.score[r];
        int node = B.node[r];
        for (int k = ln - 1; k >= 0 && node >= 0; --k) {
            P.out_tokens[o + k] = trie_tok[node];
            node = trie_parent[node];
        }
        const int head = B.times[r];
        int cnt = 0;
        for (int h = head; h >= 0; h = time_prev[h]) ++cnt;
        int k = cnt - 1;
        for (int h = head; h >= 0 && k >= 0; h = time_prev[h], --k)
            if (k < P.max_len) P.out_times[o + k] = time_t[h];
        for (int z = cnt; z < ln; ++z) P.out_times[o + z] = -1;
    } else if (tid < beam) {
        P.out_lens[utt * beam + tid] = 0;
        P.out_scores[utt * beam + tid] = neg_inf();
    }
}

}  // namespace

size_t prefix_beam_workspace_bytes(int batch, int beam, int max_len) {
    return (size_t)batch * 6 * (size_t)max_len * beam * sizeof(int) + 256;
}

int ctc_prefix_beam_search(const PrefixBeamArgs& a, cudaStream_t stream) {
    if (a.batch <= 0) return WB_OK;
    WB_REQUIRE(a.beam >= 1 && a.beam <= MAXB, WB_ERR_UNSUPPORTED, "prefix beam search: beam %d not in [1,%d]", a.beam, MAXB);
    WB_REQUIRE(a.topk >= a.beam, WB_ERR_BAD_ARG, "prefix beam search: topk %d < beam %d", a.topk, a.beam);
    WB_REQUIRE(a.workspace_bytes >= prefix_beam_workspace_bytes(a.batch, a.beam, a.max_len), WB_ERR_WORKSPACE,
               "prefix beam search: workspace too small");
    PbDev P;
    P.topk_val = a.topk_val;
    P.topk_idx = a.topk_idx;
    P.topk = a.topk;
    P.seq_start = a.seq_start;
    P.seq_len = a.seq_len;
    P.beam = a.beam;
    P.blank_id = a.blank_id;
    P.max_len = a.max_len;
    P.out_tokens = a.out_tokens;
    P.out_times = a.out_times;
    P.out_lens = a.out_lens;
    P.out_scores = a.out_scores;
    P.out_nhyp = a.out_nhyp;
    P.pool = reinterpret_cast<int*>(a.workspace);
    P.cg_nodes = a.cg_nodes;
    P.cg_child_off = a.cg_child_off;
    P.cg_child_tok = a.cg_child_tok;
    P.cg_child_node = a.cg_child_node;
    P.cg_fail = a.cg_fail;
    P.cg_token = a.cg_token;
    P.cg_node_score = a.cg_node_score;
    P.cg_token_score = a.cg_token_score;
    P.cg_output_score = a.cg_output_score;
    WB_REQUIRE(a.cg_nodes == 0 || (a.cg_child_off && a.cg_fail && a.cg_token && a.cg_node_score && a.cg_token_score &&
                                   a.cg_output_score),
               WB_ERR_BAD_ARG, "prefix beam search: incomplete context graph");
    WB_REQUIRE((long long)a.max_len * a.beam < 2147483647LL / 8, WB_ERR_UNSUPPORTED, "prefix beam search: pool too large");
    ProfScope _ps(PT_PREFIX_BEAM, stream, 0.0);
    prefix_beam_kernel<<<a.batch, PB_THREADS, 0, stream>>>(P);
    count_launch();
    WB_CHECK_LAUNCH();
    return WB_OK;
}

}  // namespace wb
