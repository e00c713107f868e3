// Beam-tree kernels: state reset, forward-batch construction, candidate merge ("select") and
// kernel (b), the beam-tree verify.  All tiny, latency-bound, single-CTA kernels that keep the search
// state on the device so a round needs exactly one 4-byte device->host read (n_matches).
//
// Reference (file:line relative to /root/reference/code):
//   select : the global part of `beam_scores.view(-1).topk(beam_size)`, `// V`, `% V`, the finite filter and
//            the construction of beam_sequence / tree-mask rows / position_ids   beamSD.py:69-91   (G3-G5)
//   verify : the greedy branch of `verify` -- level walk, hit matching, carried scores, accepted
//            length, surviving beam indices, next-round inputs and KV truncation   beamSD.py:278-330,
//            370-445                                                               (G6, G8)
// What the reference does with `.tolist()`, Python `in`, `torch.where` loops and dense masks is done
// here with warp shuffles over <= 64 candidates lists and 512-bit visibility masks.
#include "beam.cuh"

namespace atspeed {

#define L_IDX(l, i) ((l) * MAX_BEAMS + (i))

__global__ void tree_begin_kernel(TreeDev t, BatchDev b, const int* __restrict__ prompt, int P) {
    const int tid = threadIdx.x;
    if (tid < SC_COUNT) t.scal[tid] = 0;
    if (tid < MAX_LEVELS) t.cnt[tid] = tid == 0 ? 1 : 0;
    __syncthreads();
    if (tid == 0) {
        t.scal[SC_P] = P;
        t.scal[SC_FIRST] = 1;
        t.tok[0] = -1; t.parent[0] = -1; t.node[0] = 0; t.slot[0] = -1; t.score[0] = 0.f;
    }
    if (tid < MAX_NEW) t.gen[tid] = 0;
    if (tid < VIS_WORDS) t.vis[tid] = 0u;
    for (int i = tid; i < P; i += blockDim.x) {
        b.tok[i] = prompt[i]; b.pos[i] = i; b.slot[i] = i; b.prefix_len[i] = i + 1;
        for (int w = 0; w < VIS_WORDS; ++w) b.vis[i * VIS_WORDS + w] = 0u;
    }
}

int tree_begin(const TreeDev& t, const BatchDev& b, const int* prompt, int P, cudaStream_t st) {
    tree_begin_kernel<<<1, 256, 0, st>>>(t, b, prompt, P);
    ATS_LAUNCH_CHECK();
    return ATS_OK;
}

// ---------------------------------------------------------------------------------------------
// forward batch = [prompt] [missing] [levels l_from..l_to]; logits rows = [root row] [levels rows_from..l_to]
// ---------------------------------------------------------------------------------------------
__global__ void tree_build_batch_kernel(TreeDev t, BatchDev b, TreeGeom g, BatchPlan plan, const int* __restrict__ prompt,
                                        int P, int T_cap, int R_cap) {
    const int gen0 = t.scal[SC_GEN0];
    const int miss_n = t.scal[SC_MISS];
    const int trash0 = g.tree_slot(P, MAX_LEVELS, 0);    // K trash slots after the last level, for padded entries
    const int n_prompt = plan.with_prompt ? P : 0;
    const int n_miss = plan.with_missing ? g.K : 0;
    for (int x = threadIdx.x; x < T_cap; x += blockDim.x) {
        int tok = 0, pos = 0, slot = 0, prefix = 0;
        uint32_t vis[VIS_WORDS];
#pragma unroll
        for (int w = 0; w < VIS_WORDS; ++w) vis[w] = 0u;
        int y = x - n_prompt;
        if (x < n_prompt) {                               // causal prompt token (the batch arrays are reused by every forward)
            tok = prompt[x]; pos = x; slot = x; prefix = x + 1;
        } else if (y < n_miss) {
            if (y < miss_n) {
                tok = t.miss_tok[y]; pos = t.miss_pos[y]; slot = t.miss_slot[y]; prefix = P;
#pragma unroll
                for (int w = 0; w < VIS_WORDS; ++w) vis[w] = t.miss_vis[y * VIS_WORDS + w];
            } else {
                slot = trash0 + y;
                vis[(slot - P) >> 5] = 1u << ((slot - P) & 31);
            }
        } else {
            y -= n_miss;
            int l = plan.l_from;
            for (; l <= plan.l_to; ++l) {
                const int cap = l == 0 ? g.K : plan.width;
                if (y < cap) break;
                y -= cap;
            }
            if (l > plan.l_to) { slot = trash0; vis[(slot - P) >> 5] = 1u << ((slot - P) & 31); }
            else if (y < t.cnt[l]) {
                tok = t.tok[L_IDX(l, y)]; pos = P - 1 + gen0 + l; slot = t.slot[L_IDX(l, y)]; prefix = P;
#pragma unroll
                for (int w = 0; w < VIS_WORDS; ++w) vis[w] = t.vis[L_IDX(l, y) * VIS_WORDS + w];
            } else {
                slot = g.tree_slot(P, l, y);              // the slot this entry would own: nobody else reads it
                vis[(slot - P) >> 5] = 1u << ((slot - P) & 31);
            }
        }
        b.tok[x] = tok; b.pos[x] = pos; b.slot[x] = slot; b.prefix_len[x] = prefix;
#pragma unroll
        for (int w = 0; w < VIS_WORDS; ++w) b.vis[x * VIS_WORDS + w] = vis[w];
    }
    // logits rows
    for (int r = threadIdx.x; r < R_cap; r += blockDim.x) {
        int idx = 0, node = -1;
        int y = r;
        if (plan.root_row) {
            if (y == 0) { idx = P - 1; node = t.node[0]; y = -1; }
            else y -= 1;
        }
        if (y >= 0) {
            int base = n_prompt + n_miss;
            for (int l = plan.l_from; l < plan.rows_from; ++l) base += l == 0 ? g.K : plan.width;
            int l = plan.rows_from;
            for (; l <= plan.l_to; ++l) {
                const int cap = l == 0 ? g.K : plan.width;
                if (y < cap) break;
                y -= cap; base += cap;
            }
            if (l <= plan.l_to) { idx = base + y; node = y < t.cnt[l] ? t.node[L_IDX(l, y)] : -1; }
        }
        b.rows_idx[r] = idx;
        b.row_node[r] = node;
    }
}

int tree_build_batch(const TreeDev& t, const BatchDev& b, const TreeGeom& g, const BatchPlan& plan, const int* prompt,
                     int P, int T_cap, int R_cap, cudaStream_t st) {
    tree_build_batch_kernel<<<1, 256, 0, st>>>(t, b, g, plan, prompt, P, T_cap, R_cap);
    ATS_LAUNCH_CHECK();
    return ATS_OK;
}

// ---------------------------------------------------------------------------------------------
// k-way merge of per-row candidate lists (each sorted by logp desc, token asc) by one warp.
// Row j (< n_rows <= 64) contributes cand[(row_of[j]) * B + h] + parent[j]; result in rank order.
// ---------------------------------------------------------------------------------------------
struct Pick { int j, tok, edge; float score; };

__device__ __forceinline__ int warp_merge(int n_rows, const int* row_of, const float* parent, int B, int V,
                                          const int* __restrict__ cand_tok, const int* __restrict__ cand_edge,
                                          const float* __restrict__ cand_logp, const int* __restrict__ cand_cnt,
                                          int want, Pick* out /* shared */) {
    const int lane = threadIdx.x & 31;
    int head[2] = {0, 0}, cnt[2] = {0, 0}, row[2] = {0, 0};
    float par[2] = {0.f, 0.f};
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int j = lane + 32 * u;
        if (j < n_rows) { row[u] = row_of[j]; cnt[u] = cand_cnt[row[u]]; par[u] = parent[j]; }
    }
    int n_out = 0;
    for (; n_out < want; ++n_out) {
        unsigned long long best = 0ull;
        int bu = 0;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (head[u] < cnt[u]) {
                const int c = row[u] * B + head[u];
                const unsigned long long k =
                    rank_key(cand_logp[c] + par[u], static_cast<uint32_t>((lane + 32 * u) * V + cand_tok[c]));
                if (k > best) { best = k; bu = u; }
            }
        }
        const unsigned long long top = warp_max_u64(best);
        if (top == 0ull) break;
        if (best == top) {     // keys are unique, exactly one lane wins
            const int c = row[bu] * B + head[bu];
            out[n_out].j = lane + 32 * bu;
            out[n_out].tok = cand_tok[c];
            out[n_out].edge = cand_edge[c];
            out[n_out].score = cand_logp[c] + par[bu];
            head[bu]++;
        }
        __syncwarp();
    }
    return n_out;
}

__global__ void __launch_bounds__(64)
tree_select_kernel(TreeDev t, TreeGeom g, TrieCSR trie, int level, int row0, int B, const int* __restrict__ cand_tok,
                   const int* __restrict__ cand_edge, const float* __restrict__ cand_logp,
                   const int* __restrict__ cand_cnt, int width, int P) {
    __shared__ Pick picks[MAX_BEAMS];
    __shared__ int row_of[MAX_BEAMS];
    __shared__ float parent[MAX_BEAMS];
    __shared__ int s_n;
    const int n_rows = t.cnt[level];
    const int gen0 = t.scal[SC_GEN0];
    for (int j = threadIdx.x; j < n_rows; j += blockDim.x) { row_of[j] = row0 + j; parent[j] = t.score[L_IDX(level, j)]; }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int n = warp_merge(n_rows, row_of, parent, B, g.V, cand_tok, cand_edge, cand_logp, cand_cnt, width, picks);
        if (threadIdx.x == 0) s_n = n;
    }
    __syncthreads();
    const int n = s_n, nl = level + 1;
    for (int p = threadIdx.x; p < n; p += blockDim.x) {
        const Pick pk = picks[p];
        const int slot = g.tree_slot(P, nl, p);
        t.tok[L_IDX(nl, p)] = pk.tok;
        t.parent[L_IDX(nl, p)] = pk.j;
        t.score[L_IDX(nl, p)] = pk.score;
        t.node[L_IDX(nl, p)] = trie.child_node[pk.edge];
        t.slot[L_IDX(nl, p)] = slot;
        for (int k = 0; k < MAX_NEW; ++k) t.gen[L_IDX(nl, p) * MAX_NEW + k] = t.gen[L_IDX(level, pk.j) * MAX_NEW + k];
        if (gen0 + level < MAX_NEW) t.gen[L_IDX(nl, p) * MAX_NEW + gen0 + level] = pk.tok;
        for (int w = 0; w < VIS_WORDS; ++w) t.vis[L_IDX(nl, p) * VIS_WORDS + w] = t.vis[L_IDX(level, pk.j) * VIS_WORDS + w];
        t.vis[L_IDX(nl, p) * VIS_WORDS + ((slot - P) >> 5)] |= 1u << ((slot - P) & 31);
    }
    if (threadIdx.x == 0) { t.cnt[nl] = n; t.scal[SC_RESULT] = nl; }
}

int tree_select(const TreeDev& t, const TreeGeom& g, const TrieCSR& trie, int level, int row0, int B,
                const int* cand_tok, const int* cand_edge, const float* cand_logp, const int* cand_cnt, int width,
                int P, cudaStream_t st) {
    ATS_CHECK_ARG(level >= 0 && level + 1 < MAX_LEVELS && width >= 1 && width <= MAX_BEAMS, "select: level=%d width=%d",
                  level, width);
    tree_select_kernel<<<1, 64, 0, st>>>(t, g, trie, level, row0, B, cand_tok, cand_edge, cand_logp, cand_cnt, width, P);
    ATS_LAUNCH_CHECK();
    return ATS_OK;
}

// ---------------------------------------------------------------------------------------------
// kernel (b): AtSpeed-S strict top-K verify over the whole draft tree in one launch
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64)
tree_verify_strict_kernel(TreeDev t, TreeGeom g, TrieCSR trie, int draft_len, int root_rows,
                          const int* __restrict__ cand_tok, const int* __restrict__ cand_edge,
                          const float* __restrict__ cand_logp, const int* __restrict__ cand_cnt, int P) {
    __shared__ Pick picks[MAX_BEAMS];
    __shared__ int row_of[MAX_BEAMS];
    __shared__ float cur_score[MAX_BEAMS];
    __shared__ int cur_idx[MAX_BEAMS];
    __shared__ int pick_pos[MAX_BEAMS];
    __shared__ int stage_idx[MAX_K];
    __shared__ float stage_score[MAX_K];
    __shared__ int s_npk, s_hits, s_cur_n;
    __shared__ unsigned char used[MAX_LEVELS][MAX_BEAMS];
    __shared__ int newslot[MAX_LEVELS][MAX_BEAMS];
    __shared__ uint32_t newvis[MAX_LEVELS][MAX_K][VIS_WORDS];   // indexed by compact id
    __shared__ int compact[MAX_LEVELS][MAX_BEAMS];              // (level, idx) -> compact id among used nodes of the level
    __shared__ int r_gen[MAX_K][MAX_NEW];
    __shared__ uint32_t r_vis[MAX_K][VIS_WORDS];
    __shared__ int r_node[MAX_K];

    const int tid = threadIdx.x, K = g.K, N = g.N, V = g.V;
    const int gen0 = t.scal[SC_GEN0], first = t.scal[SC_FIRST], acc0 = t.scal[SC_ACC];

    // ---- level walk (beamSD.py:278-380) ----
    if (tid == 0) s_cur_n = t.cnt[0];
    for (int j = tid; j < t.cnt[0]; j += blockDim.x) { cur_idx[j] = j; cur_score[j] = t.score[L_IDX(0, j)]; }
    __syncthreads();
    int m = 0, lvl = 0;
    for (lvl = 0; lvl <= draft_len; ++lvl) {
        const int rowbase = lvl == 0 ? 0 : root_rows + (lvl - 1) * N;
        const int cur_n = s_cur_n;
        for (int j = tid; j < cur_n; j += blockDim.x) row_of[j] = rowbase + cur_idx[j];
        __syncthreads();
        if (tid < 32) {
            const int n = warp_merge(cur_n, row_of, cur_score, K, V, cand_tok, cand_edge, cand_logp, cand_cnt, K, picks);
            if (tid == 0) { s_npk = n; t.tr_npick[lvl] = n; }
        }
        __syncthreads();
        const int npk = s_npk;
        if (tid == 0) s_hits = 0;
        __syncthreads();
        // match the target's picks against the draft's next level (pair = (draft beam index, token))
        const int n_next = lvl < draft_len ? t.cnt[lvl + 1] : 0;
        for (int p = tid; p < npk; p += blockDim.x) {
            const int par = cur_idx[picks[p].j];
            int pos = -1;
            for (int qd = 0; qd < n_next; ++qd)
                if (t.parent[L_IDX(lvl + 1, qd)] == par && t.tok[L_IDX(lvl + 1, qd)] == picks[p].tok) { pos = qd; break; }
            pick_pos[p] = pos;
            if (pos >= 0) atomicAdd(&s_hits, 1);
            t.tr_pick_parent[lvl * MAX_K + p] = par;
            t.tr_pick_tok[lvl * MAX_K + p] = picks[p].tok;
            t.tr_pick_score[lvl * MAX_K + p] = picks[p].score;
            t.tr_hit_pos[lvl * MAX_K + p] = pos;
        }
        __syncthreads();
        if (lvl == draft_len || s_hits != K) break;          // bonus level reached, or the level is rejected
        // accepted: carry the K hit beams in draft-position order with the TARGET's scores
        for (int p = tid; p < npk; p += blockDim.x) {
            int rank = 0;
            for (int o = 0; o < npk; ++o) rank += pick_pos[o] < pick_pos[p];
            stage_idx[rank] = pick_pos[p];
            stage_score[rank] = picks[p].score;
        }
        __syncthreads();
        for (int j = tid; j < K; j += blockDim.x) { cur_idx[j] = stage_idx[j]; cur_score[j] = stage_score[j]; }
        if (tid == 0) s_cur_n = K;
        ++m;
        __syncthreads();
    }
    const int npk = s_npk;                                   // final beams = picks of level m, in score order

    // ---- which tree nodes do the survivors descend from? ----
    for (int i = tid; i < MAX_LEVELS * MAX_BEAMS; i += blockDim.x) (&used[0][0])[i] = 0;
    __syncthreads();
    if (tid < npk) {
        int a = cur_idx[picks[tid].j];
        for (int l = m; l >= 0; --l) { used[l][a] = 1; a = l > 0 ? t.parent[L_IDX(l, a)] : a; }
    }
    __syncthreads();
    // compact slots in the accepted region, level by level (first round: level 0 is the prompt itself)
    if (tid == 0) {
        int n_new = 0;
        for (int l = first ? 1 : 0; l <= m; ++l) {
            int cid = 0;
            for (int qd = 0; qd < t.cnt[l]; ++qd)
                if (used[l][qd]) {
                    newslot[l][qd] = P + acc0 + n_new;
                    compact[l][qd] = cid++;
                    t.gather_src[n_new] = t.slot[L_IDX(l, qd)];
                    t.gather_dst[n_new] = P + acc0 + n_new;
                    ++n_new;
                }
        }
        if (first) { compact[0][0] = 0; newslot[0][0] = -1; }
        t.scal[SC_GATHER] = n_new;
        t.scal[SC_ACC] = acc0 + n_new;
    }
    __syncthreads();
    // visibility of the kept nodes in the new (compacted) slot numbering
    const int acc_words = (g.A_cap + 31) >> 5;
    for (int l = 0; l <= m; ++l) {
        for (int qd = tid; qd < t.cnt[l]; qd += blockDim.x) {
            if (!used[l][qd]) continue;
            const int cid = compact[l][qd];
            for (int w = 0; w < VIS_WORDS; ++w) {
                uint32_t v;
                if (l == 0) {
                    v = first ? 0u : t.vis[L_IDX(0, qd) * VIS_WORDS + w];
                    if (w >= acc_words) v = 0u;
                    else if (w == acc_words - 1 && (g.A_cap & 31)) v &= (1u << (g.A_cap & 31)) - 1u;
                } else {
                    v = newvis[l - 1][compact[l - 1][t.parent[L_IDX(l, qd)]]][w];
                }
                newvis[l][cid][w] = v;
            }
            if (!(first && l == 0)) {
                const int bit = newslot[l][qd] - P;
                newvis[l][cid][bit >> 5] |= 1u << (bit & 31);
            }
        }
        __syncthreads();
    }
    // draft tokens without KV: the kept nodes of the last level when every level was accepted
    if (tid == 0) t.scal[SC_MISS] = 0;
    __syncthreads();
    if (m == draft_len && draft_len > 0) {
        for (int qd = tid; qd < t.cnt[m]; qd += blockDim.x) {
            if (!used[m][qd]) continue;
            const int cid = compact[m][qd];
            t.miss_tok[cid] = t.tok[L_IDX(m, qd)];
            t.miss_pos[cid] = P - 1 + gen0 + m;
            t.miss_slot[cid] = newslot[m][qd];
            for (int w = 0; w < VIS_WORDS; ++w) t.miss_vis[cid * VIS_WORDS + w] = newvis[m][cid][w];
            atomicAdd(&t.scal[SC_MISS], 1);
        }
    }
    // ---- next round's roots (staged in shared memory: they overwrite level 0) ----
    if (tid < npk) {
        const Pick pk = picks[tid];
        const int a = cur_idx[pk.j];
        for (int k = 0; k < MAX_NEW; ++k) r_gen[tid][k] = t.gen[L_IDX(m, a) * MAX_NEW + k];
        if (gen0 + m < MAX_NEW) r_gen[tid][gen0 + m] = pk.tok;
        const int slot = g.tree_slot(P, 0, tid);
        for (int w = 0; w < VIS_WORDS; ++w) r_vis[tid][w] = newvis[m][compact[m][a]][w];
        r_vis[tid][(slot - P) >> 5] |= 1u << ((slot - P) & 31);
        r_node[tid] = trie.child_node[pk.edge];
    }
    __syncthreads();
    if (tid < npk) {
        const Pick pk = picks[tid];
        t.tok[L_IDX(0, tid)] = pk.tok;
        t.parent[L_IDX(0, tid)] = cur_idx[pk.j];             // index in the level the beam descends from (level m)
        t.score[L_IDX(0, tid)] = pk.score;
        t.node[L_IDX(0, tid)] = r_node[tid];
        t.slot[L_IDX(0, tid)] = g.tree_slot(P, 0, tid);
        for (int k = 0; k < MAX_NEW; ++k) t.gen[L_IDX(0, tid) * MAX_NEW + k] = r_gen[tid][k];
        for (int w = 0; w < VIS_WORDS; ++w) t.vis[L_IDX(0, tid) * VIS_WORDS + w] = r_vis[tid][w];
    }
    if (tid == 0) {
        t.cnt[0] = npk;
        for (int l = 1; l < MAX_LEVELS; ++l) t.cnt[l] = 0;
        t.scal[SC_NMATCH] = m;
        t.scal[SC_GEN0] = gen0 + m + 1;
        t.scal[SC_FIRST] = 0;
        t.scal[SC_RESULT] = 0;
    }
}

int tree_verify_strict(const TreeDev& t, const TreeGeom& g, const TrieCSR& trie, int draft_len, int root_rows,
                       const int* cand_tok, const int* cand_edge, const float* cand_logp, const int* cand_cnt, int P,
                       cudaStream_t st) {
    ATS_CHECK_ARG(draft_len >= 1 && draft_len + 1 < MAX_LEVELS, "verify: draft_len=%d", draft_len);
    ATS_CHECK_ARG(g.K <= MAX_K, "verify: K=%d > %d", g.K, MAX_K);
    tree_verify_strict_kernel<<<1, 64, 0, st>>>(t, g, trie, draft_len, root_rows, cand_tok, cand_edge, cand_logp,
                                                cand_cnt, P);
    ATS_LAUNCH_CHECK();
    return ATS_OK;
}

}  // namespace atspeed
