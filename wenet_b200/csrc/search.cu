// CTC prefix beam search on the GPU: one CTA per utterance, frames sequential, all (token, prefix)
// pairs of a frame in parallel.  Bit-level restatement of the reference's host loop
//   wenet/models/transformer/search.py:127-249 (ctc_prefix_beam_search, PrefixScore :64-106)
//   wenet/utils/common.py:302-310 (log_add, Python float == IEEE double)
// including: double-precision scores, Viterbi scores / token times (times_s / times_ns /
// cur_token_prob rules of :166-219), dict insertion order as the tie-break of the stable
// `sorted(..., reverse=True)` (:222-225), first-max-wins comparisons.
//
// Data structures
//   * prefixes are nodes of a trie (parent, token) kept in a per-utterance pool in HBM; identity of
//     a prefix inside a frame is (64-bit rolling hash, length, last token)
//   * times lists are persistent (immutable) linked lists (prev, frame) in a second pool, so
//     `list.copy()` / `append` / `[-1] = t` of the reference are O(1) pointer operations
//   * the beam (<= 16 entries) and the <= beam + beam^2 candidates of a frame live in shared memory.
// Each "unchanged prefix" candidate receives at most three updates per frame (blank, repeat,
// extension-of-its-parent) — they are replayed in the reference's loop order (token-major,
// prefix-minor) so every floating-point operation happens in the same order as on the host.
#include "common.cuh"
#include "kernels.h"
#include <limits.h>
#include <math_constants.h>
#include "softplus_table.h"

namespace wb {

namespace {

constexpr int MAXB = 16;
constexpr int NCAND = MAXB + MAXB * MAXB;
constexpr int PB_WARPS = 8;             // utterances per CTA: one warp each, every barrier is a __syncwarp, so the
                                        // whole batch occupies only batch/8 SMs and other streams keep the rest
constexpr int PB_THREADS = 32 * PB_WARPS;
constexpr int MAXPL = (NCAND + 31) / 32;  // candidates per lane in the selection step

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
    double coef[WB_SOFTPLUS_DEG + 1];
#pragma unroll
    for (int j = 0; j <= WB_SOFTPLUS_DEG; ++j) coef[j] = __ldg(c + j);
    double v = coef[WB_SOFTPLUS_DEG];
#pragma unroll
    for (int j = WB_SOFTPLUS_DEG - 1; j >= 0; --j) v = fma(v, t, coef[j]);
    return v;
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

__device__ __forceinline__ uint64_t mix_hash(uint64_t h, int tok) {
    uint64_t z = h ^ ((uint64_t)(uint32_t)(tok + 1) * 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

struct Beam {
    double s[MAXB], ns[MAXB], vs[MAXB], vns[MAXB];
    double score[MAXB], vit[MAXB];
    uint64_t hash[MAXB];
    int len[MAXB], last[MAXB], node[MAXB], ts[MAXB], tns[MAXB], times[MAXB];
    int n;
};

struct Cand {
    double s[NCAND], ns[NCAND], vs[NCAND], vns[NCAND], total[NCAND];
    uint64_t hash[NCAND];
    int len[NCAND], last[NCAND];
    int node[NCAND];       // existing trie node (unchanged prefix) or parent node (extension)
    int new_tok[NCAND];    // >= 0: extension by this token (needs a new trie node)
    int ts[NCAND];         // times_s head
    int tns[NCAND];        // times_ns head if tns_new == 0, else prev pointer of the node to create
    int tns_new[NCAND];    // 1: times_ns = list(tns) + [t]
    int first[NCAND];      // first-touch sequence number == dict insertion order
    int valid[NCAND];
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
    int* pool;  // per utterance: [4][max_len * beam] ints: trie parent, trie token, time prev, time frame
};

__device__ __forceinline__ bool cand_better(double ta, int fa, double tb, int fb) {
    return ta > tb || (ta == tb && fa < fb);
}

struct WarpState {
    Beam Bs[2];
    Cand C;
    float tk_val[MAXB];
    int tk_idx[MAXB];
    int dest[MAXB * MAXB];
    int rank_slot[MAXB];
    int vlist[NCAND];
};

__global__ void __launch_bounds__(PB_THREADS)
prefix_beam_kernel(PbDev P, int batch) {
    extern __shared__ __align__(16) uint8_t pb_smem[];
    const int utt = blockIdx.x * PB_WARPS + (threadIdx.x >> 5);
    if (utt >= batch) return;  // whole warp; no block-level barriers are used anywhere below
    WarpState& W = reinterpret_cast<WarpState*>(pb_smem)[threadIdx.x >> 5];
    Beam* Bs = W.Bs;
    Cand& C = W.C;
    float* tk_val = W.tk_val;
    int* tk_idx = W.tk_idx;
    int* dest = W.dest;
    int* rank_slot = W.rank_slot;
    int* vlist = W.vlist;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int beam = P.beam;
    const int T = P.seq_len[utt];
    const long long f0 = P.seq_start[utt];
    const long long pool_n = (long long)P.max_len * beam;
    int* trie_parent = P.pool + (long long)utt * 4 * pool_n;
    int* trie_tok = trie_parent + pool_n;
    int* time_prev = trie_tok + pool_n;
    int* time_t = time_prev + pool_n;
    const int ncs = MAXB + beam * MAXB;  // candidate slots in use

    for (int c = lane; c < NCAND; c += 32) C.valid[c] = 0;
    if (lane == 0) {
        Beam& B = Bs[0];
        B.n = 1;
        B.s[0] = 0.0;
        B.ns[0] = neg_inf();
        B.vs[0] = 0.0;
        B.vns[0] = 0.0;
        B.score[0] = 0.0;   // log_add(0, -inf)
        B.vit[0] = 0.0;     // v_s > v_ns is false -> v_ns
        B.times[0] = -1;
        B.hash[0] = 0x1234567ull;
        B.len[0] = 0;
        B.last[0] = -1;
        B.node[0] = -1;
        B.ts[0] = -1;
        B.tns[0] = -1;
    }
    int cur = 0;
    float pf_val = 0.f;
    int pf_idx = 0;
    if (lane < beam && T > 0) {
        pf_val = P.topk_val[f0 * P.topk + lane];
        pf_idx = P.topk_idx[f0 * P.topk + lane];
    }
    __syncwarp();

    for (int t = 0; t < T; ++t) {
        if (lane < beam) {
            tk_val[lane] = pf_val;
            tk_idx[lane] = pf_idx;
        }
        __syncwarp();
        if (lane < beam && t + 1 < T) {  // prefetch the next frame's top-k behind this frame's work
            pf_val = P.topk_val[(f0 + t + 1) * P.topk + lane];
            pf_idx = P.topk_idx[(f0 + t + 1) * P.topk + lane];
        }
        Beam& B = Bs[cur];
        const int nb = B.n;

        // ---- extensions: (token ui, prefix pi) pairs strided over the lanes ----
        for (int pr = lane; pr < beam * nb; pr += 32) {
            const int ui = pr / nb, pi = pr - ui * nb;
            const int u = tk_idx[ui];
            int d = -1;
            if (u != P.blank_id) {
                const uint64_t h = mix_hash(B.hash[pi], u);
                const int ln = B.len[pi] + 1;
                d = MAXB + ui * MAXB + pi;
                for (int q = 0; q < nb; ++q)
                    if (B.last[q] == u && B.len[q] == ln && B.hash[q] == h) d = q;
                if (d >= MAXB) {
                    const double prob = (double)tk_val[ui];
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
                    C.hash[d] = h;
                    C.len[d] = ln;
                    C.last[d] = u;
                    C.node[d] = B.node[pi];
                    C.new_tok[d] = u;
                    C.first[d] = (ui * nb + pi) * 2 + (rep ? 1 : 0);
                    C.total[d] = nsv;  // log_add(-inf, ns)
                    C.valid[d] = t + 1;   // frame stamp: no per-frame reset of the flags
                }
            }
            dest[ui * MAXB + pi] = d;
        }
        __syncwarp();

        // ---- unchanged prefixes: lane q replays its (<= 3) updates in the reference's loop order ----
        if (lane < nb) {
            const int q = lane;
            const int lastq = B.last[q];
            int ui_blank = -1, ui_last = -1;
            for (int ui = 0; ui < beam; ++ui) {
                const int u = tk_idx[ui];
                if (u == P.blank_id) ui_blank = ui;
                else if (u == lastq) ui_last = ui;
            }
            double s = neg_inf(), ns = neg_inf(), vs = neg_inf(), vns = neg_inf();
            int ts = -1, tns = -1, tns_new = 0, first = INT_MAX, any = 0;
            if (ui_blank >= 0) {
                const double prob = (double)tk_val[ui_blank];
                s = B.score[q] + prob;          // log_add(-inf, x)
                vs = B.vit[q] + prob;
                ts = B.times[q];
                first = (ui_blank * nb + q) * 2;
                any = 1;
            }
            if (ui_last >= 0) {
                const double prob = (double)tk_val[ui_last];
                int pe = -1;  // parent prefix whose extension by last(q) lands on q
                for (int pi = 0; pi < nb; ++pi)
                    if (pi != q && dest[ui_last * MAXB + pi] == q) pe = pi;
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
                C.hash[q] = B.hash[q];
                C.len[q] = B.len[q];
                C.last[q] = lastq;
                C.node[q] = B.node[q];
                C.new_tok[q] = -1;
                C.first[q] = first;
                C.total[q] = log_add2(s, ns);
                C.valid[q] = t + 1;
            }
        }
        __syncwarp();

        // ---- second beam prune: stable sort by total desc == top-beam by (total, insertion order) ----
        int nvalid = 0;
        for (int base = 0; base < ncs; base += 32) {
            const int c = base + lane;
            const bool v = (c < ncs) && (C.valid[c] == t + 1);
            const unsigned m = __ballot_sync(0xffffffffu, v);
            if (v) vlist[nvalid + __popc(m & lt_mask)] = c;
            nvalid += __popc(m);
        }
        __syncwarp();
        // Exact top-`beam` of <= beam + beam^2 candidates without serial arg-max rounds:
        //  (1) every lane reduces its own <= MAXPL candidates to a local best; the `beam`-th best of the 32 local
        //      bests is a lower bound of the true `beam`-th best total, so everything strictly below it is pruned
        //      (ranks among lanes by 31 independent shuffle rotations);
        //  (2) the survivors (>= beam, typically 10-15) are compacted one per lane and ranked the same way.
        //  Order = the reference's stable descending sort: (total desc, dict insertion order asc).
        const int kmax = (nvalid + 31) >> 5;  // warp-uniform
        double my_tot[MAXPL];
        int my_first[MAXPL], my_slot[MAXPL];
        double lt = neg_inf();
        int lf = INT_MAX;
#pragma unroll
        for (int k = 0; k < MAXPL; ++k) {
            my_slot[k] = -1;
            my_tot[k] = neg_inf();
            my_first[k] = INT_MAX;
            if (k < kmax) {
                const int i = lane + 32 * k;
                if (i < nvalid) {
                    const int c = vlist[i];
                    my_slot[k] = c;
                    my_tot[k] = C.total[c];
                    my_first[k] = C.first[c];
                    if (cand_better(my_tot[k], my_first[k], lt, lf)) {
                        lt = my_tot[k];
                        lf = my_first[k];
                    }
                }
            }
        }
        double thr_t = neg_inf();
        int thr_f = INT_MAX;  // default: nothing is pruned
        if (nvalid > 32) {
            int rk = 0;
#pragma unroll
            for (int r = 1; r < 32; ++r) {
                const double ot = __shfl_sync(0xffffffffu, lt, (lane + r) & 31);
                const int of = __shfl_sync(0xffffffffu, lf, (lane + r) & 31);
                rk += cand_better(ot, of, lt, lf) ? 1 : 0;
            }
            const unsigned m = __ballot_sync(0xffffffffu, rk == beam - 1);  // all 32 lanes hold a candidate here
            if (m) {
                const int src = __ffs(m) - 1;
                thr_t = __shfl_sync(0xffffffffu, lt, src);
                thr_f = __shfl_sync(0xffffffffu, lf, src);
            }
        }
        __syncwarp();  // vlist fully consumed -> reuse it for the survivor list
        int ns = 0;
#pragma unroll
        for (int k = 0; k < MAXPL; ++k) {
            if (k < kmax) {
                const bool sv = my_slot[k] >= 0 && !cand_better(thr_t, thr_f, my_tot[k], my_first[k]);
                const unsigned m = __ballot_sync(0xffffffffu, sv);
                if (sv) vlist[ns + __popc(m & lt_mask)] = my_slot[k];
                ns += __popc(m);
            }
        }
        __syncwarp();
        const int nnew = nvalid < beam ? nvalid : beam;
        if (ns <= 32) {
            const int c = lane < ns ? vlist[lane] : -1;
            const double t_ = c >= 0 ? C.total[c] : neg_inf();
            const int f_ = c >= 0 ? C.first[c] : INT_MAX;
            int rk = 0;
#pragma unroll
            for (int r = 1; r < 32; ++r) {
                const double ot = __shfl_sync(0xffffffffu, t_, (lane + r) & 31);
                const int of = __shfl_sync(0xffffffffu, f_, (lane + r) & 31);
                rk += cand_better(ot, of, t_, f_) ? 1 : 0;   // empty lanes carry (-inf, INT_MAX): never better
            }
            if (c >= 0 && rk < beam) rank_slot[rk] = c;
        } else {
            // rare fallback (many candidates tie with / exceed the bound): serial arg-max rounds over the survivors
            const int kmax2 = (ns + 31) >> 5;
#pragma unroll
            for (int k = 0; k < MAXPL; ++k) {
                my_slot[k] = -1;
                if (k < kmax2 && lane + 32 * k < ns) {
                    const int c = vlist[lane + 32 * k];
                    my_slot[k] = c;
                    my_tot[k] = C.total[c];
                    my_first[k] = C.first[c];
                }
            }
            for (int r = 0; r < nnew; ++r) {
                double bt = neg_inf();
                int bf = INT_MAX, bs = -1;
#pragma unroll
                for (int k = 0; k < MAXPL; ++k)
                    if (k < kmax2 && my_slot[k] >= 0 && cand_better(my_tot[k], my_first[k], bt, bf)) {
                        bt = my_tot[k];
                        bf = my_first[k];
                        bs = my_slot[k];
                    }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ot = __shfl_xor_sync(0xffffffffu, bt, o);
                    const int of = __shfl_xor_sync(0xffffffffu, bf, o);
                    const int os = __shfl_xor_sync(0xffffffffu, bs, o);
                    if (cand_better(ot, of, bt, bf)) {
                        bt = ot;
                        bf = of;
                        bs = os;
                    }
                }
#pragma unroll
                for (int k = 0; k < MAXPL; ++k)
                    if (k < kmax2 && my_slot[k] == bs) my_slot[k] = -1;
                if (lane == 0) rank_slot[r] = bs;
            }
        }
        __syncwarp();

        // ---- materialise the new beam ----
        Beam& NB = Bs[cur ^ 1];
        if (lane < nnew) {
            const int r = lane;
            const int c = rank_slot[r];
            const int pool_i = t * beam + r;
            const double vs = C.vs[c], vns = C.vns[c];
            NB.s[r] = C.s[c];
            NB.ns[r] = C.ns[c];
            NB.vs[r] = vs;
            NB.vns[r] = vns;
            NB.score[r] = C.total[c];
            NB.hash[r] = C.hash[c];
            NB.len[r] = C.len[c];
            NB.last[r] = C.last[c];
            if (C.new_tok[c] >= 0) {
                trie_parent[pool_i] = C.node[c];
                trie_tok[pool_i] = C.new_tok[c];
                NB.node[r] = pool_i;
            } else {
                NB.node[r] = C.node[c];
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
        if (lane == 0) NB.n = nnew;
        __syncwarp();
        cur ^= 1;
    }

    // ---- emit n-best ----
    __threadfence_block();
    __syncwarp();
    const Beam& B = Bs[cur];
    const int nb = B.n;
    if (lane == 0) P.out_nhyp[utt] = nb;
    if (lane < nb) {
        const int r = lane;
        const long long o = ((long long)utt * beam + r) * P.max_len;
        const int ln = B.len[r];
        P.out_lens[utt * beam + r] = ln;
        P.out_scores[utt * beam + r] = B.score[r];
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
    } else if (lane < beam) {
        P.out_lens[utt * beam + lane] = 0;
        P.out_scores[utt * beam + lane] = neg_inf();
    }
}

}  // namespace

size_t prefix_beam_workspace_bytes(int batch, int beam, int max_len) {
    return (size_t)batch * 4 * (size_t)max_len * beam * sizeof(int) + 256;
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
    const size_t smem = sizeof(WarpState) * PB_WARPS;
    WB_SET_MAX_DYN_SMEM(prefix_beam_kernel, smem);
    ProfScope _ps(PT_PREFIX_BEAM, stream, 0.0);
    prefix_beam_kernel<<<ceil_div(a.batch, PB_WARPS), PB_THREADS, smem, stream>>>(P, a.batch);
    count_launch();
    WB_CHECK_LAUNCH();
    return WB_OK;
}

}  // namespace wb
