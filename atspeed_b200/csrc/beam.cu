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
__device__ void tree_build_batch_body(const TreeDev& t, const BatchDev& b, const TreeGeom& g, const BatchPlan& plan,
                                      const int* __restrict__ prompt, int P, int T_cap, int R_cap, int tok0, int row0,
                                      int user) {
    const int gen0 = t.scal[SC_GEN0];
    const int miss_n = t.scal[SC_MISS];
    const int trash0 = g.tree_slot(P, MAX_LEVELS, 0);    // K trash slots after the last level, for padded entries
    const int skip = plan.with_prompt ? plan.prompt_skip : 0;
    const int n_prompt = plan.with_prompt ? P - skip : 0;
    const int n_miss = plan.with_missing ? g.K : 0;
    for (int x = threadIdx.x; x < T_cap; x += blockDim.x) {
        int tok = 0, pos = 0, slot = 0, prefix = 0;
        uint32_t vis[VIS_WORDS];
#pragma unroll
        for (int w = 0; w < VIS_WORDS; ++w) vis[w] = 0u;
        int y = x - n_prompt;
        if (x < n_prompt) {                               // causal prompt token (the batch arrays are reused by every forward)
            tok = prompt[x + skip]; pos = x + skip; slot = x + skip; prefix = x + skip + 1;
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
        const int xo = tok0 + x;
        b.tok[xo] = tok; b.pos[xo] = pos; b.slot[xo] = slot; b.prefix_len[xo] = prefix;
        if (user >= 0) b.tok_user[xo] = user;
#pragma unroll
        for (int w = 0; w < VIS_WORDS; ++w) b.vis[xo * VIS_WORDS + w] = vis[w];
    }
    // logits rows
    for (int r = threadIdx.x; r < R_cap; r += blockDim.x) {
        int idx = 0, node = -1;
        int y = r;
        if (plan.root_row) {
            if (y == 0) { idx = P - 1 - skip; node = t.node[0]; y = -1; }
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
        b.rows_idx[row0 + r] = tok0 + idx;
        b.row_node[row0 + r] = node;
    }
}

__global__ void tree_build_batch_kernel(TreeDev t, BatchDev b, TreeGeom g, BatchPlan plan, const int* __restrict__ prompt,
                                        int P, int T_cap, int R_cap) {
    tree_build_batch_body(t, b, g, plan, prompt, P, T_cap, R_cap, 0, 0, -1);
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

__device__ void tree_select_body(const TreeDev& t, const TreeGeom& g, const TrieCSR& trie, int level, int row0, int B,
                                 const int* __restrict__ cand_tok, const int* __restrict__ cand_edge,
                                 const float* __restrict__ cand_logp, const int* __restrict__ cand_cnt, int width, int P) {
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

__global__ void __launch_bounds__(64)
tree_select_kernel(TreeDev t, TreeGeom g, TrieCSR trie, int level, int row0, int B, const int* __restrict__ cand_tok,
                   const int* __restrict__ cand_edge, const float* __restrict__ cand_logp,
                   const int* __restrict__ cand_cnt, int width, int P) {
    tree_select_body(t, g, trie, level, row0, B, cand_tok, cand_edge, cand_logp, cand_cnt, width, P);
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
// Shared tail of both verify kernels: given the accepted length m and the round's final beams
// (fin_*: parent position in level m, token, trie edge, score -- shared-memory arrays of npk entries),
// mark the survivors' ancestors, emit the KV compaction lists for kernel (c), renumber visibility masks,
// list the draft's missing tokens and install the next round's roots (beamSD.py:381-445).
// ---------------------------------------------------------------------------------------------
__device__ void round_tail(const TreeDev& t, const TreeGeom& g, const TrieCSR& trie, int P, int m, int npk, int draft_len,
                           const int* fin_parent, const int* fin_tok, const int* fin_edge, const float* fin_score) {
    __shared__ unsigned char used[MAX_LEVELS][MAX_BEAMS];
    __shared__ int newslot[MAX_LEVELS][MAX_BEAMS];
    __shared__ uint32_t newvis[MAX_LEVELS][MAX_K][VIS_WORDS];   // indexed by compact id
    __shared__ int compact[MAX_LEVELS][MAX_BEAMS];              // (level, idx) -> compact id among used nodes of the level
    __shared__ int r_gen[MAX_K][MAX_NEW];
    __shared__ uint32_t r_vis[MAX_K][VIS_WORDS];
    __shared__ int r_node[MAX_K];
    const int tid = threadIdx.x;
    const int gen0 = t.scal[SC_GEN0], first = t.scal[SC_FIRST], acc0 = t.scal[SC_ACC];
    __syncthreads();

    // ---- which tree nodes do the survivors descend from? ----
    for (int i = tid; i < MAX_LEVELS * MAX_BEAMS; i += blockDim.x) (&used[0][0])[i] = 0;
    __syncthreads();
    if (tid < npk) {
        int a = fin_parent[tid];
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
        const int a = fin_parent[tid];
        for (int k = 0; k < MAX_NEW; ++k) r_gen[tid][k] = t.gen[L_IDX(m, a) * MAX_NEW + k];
        if (gen0 + m < MAX_NEW) r_gen[tid][gen0 + m] = fin_tok[tid];
        const int slot = g.tree_slot(P, 0, tid);
        for (int w = 0; w < VIS_WORDS; ++w) r_vis[tid][w] = newvis[m][compact[m][a]][w];
        r_vis[tid][(slot - P) >> 5] |= 1u << ((slot - P) & 31);
        r_node[tid] = trie.child_node[fin_edge[tid]];
    }
    __syncthreads();
    if (tid < npk) {
        t.tok[L_IDX(0, tid)] = fin_tok[tid];
        t.parent[L_IDX(0, tid)] = fin_parent[tid];           // index in the level the beam descends from (level m)
        t.score[L_IDX(0, tid)] = fin_score[tid];
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

// ---------------------------------------------------------------------------------------------
// kernel (b): AtSpeed-S strict top-K verify over the whole draft tree in one launch
// ---------------------------------------------------------------------------------------------
__device__ void tree_verify_strict_body(const TreeDev& t, const TreeGeom& g, const TrieCSR& trie, int draft_len, int root_rows,
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
    __shared__ int fin_parent[MAX_K], fin_tok[MAX_K], fin_edge[MAX_K];
    __shared__ float fin_score[MAX_K];

    const int tid = threadIdx.x, K = g.K, N = g.N, V = g.V;

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
    // final beams = picks of level m, in score order
    const int npk = s_npk;
    if (tid < npk) {
        fin_parent[tid] = cur_idx[picks[tid].j];
        fin_tok[tid] = picks[tid].tok;
        fin_edge[tid] = picks[tid].edge;
        fin_score[tid] = picks[tid].score;
    }
    __syncthreads();
    round_tail(t, g, trie, P, m, npk, draft_len, fin_parent, fin_tok, fin_edge, fin_score);
}

__global__ void __launch_bounds__(64)
tree_verify_strict_kernel(TreeDev t, TreeGeom g, TrieCSR trie, int draft_len, int root_rows,
                          const int* __restrict__ cand_tok, const int* __restrict__ cand_edge,
                          const float* __restrict__ cand_logp, const int* __restrict__ cand_cnt, int P) {
    tree_verify_strict_body(t, g, trie, draft_len, root_rows, cand_tok, cand_edge, cand_logp, cand_cnt, P);
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

// =============================================================================================
// AtSpeed-R: sampling select + relaxed (speculative-sampling) verify
// =============================================================================================
namespace atspeed {

static constexpr int SAMPLE_THREADS = 256;
static constexpr int SAMPLE_MAX_CAND = MAX_BEAMS * MAX_BEAMS;   // rows x warped candidates per row

struct LseAcc { float m, s; };
__device__ __forceinline__ LseAcc lse_merge(LseAcc a, LseAcc b) {
    if (b.m == -INFINITY) return a;
    if (a.m == -INFINITY) return b;
    LseAcc r;
    r.m = fmaxf(a.m, b.m);
    r.s = a.s * expf(a.m - r.m) + b.s * expf(b.m - r.m);
    return r;
}
// log-sum-exp over one value per (thread, iteration); `acc` holds this thread's partial
__device__ float block_lse(LseAcc acc) {
    __shared__ LseAcc red[SAMPLE_THREADS / 32];
    __shared__ float out;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        LseAcc other;
        other.m = __shfl_xor_sync(0xffffffffu, acc.m, o);
        other.s = __shfl_xor_sync(0xffffffffu, acc.s, o);
        acc = lse_merge(acc, other);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        LseAcc tot = red[0];
        for (int w = 1; w < (blockDim.x >> 5); ++w) tot = lse_merge(tot, red[w]);
        out = tot.m == -INFINITY ? -INFINITY : tot.m + logf(tot.s);
    }
    __syncthreads();
    return out;
}
__device__ float block_sum(float v) {
    __shared__ float red[SAMPLE_THREADS / 32];
    __shared__ float out;
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int w = 0; w < (blockDim.x >> 5); ++w) tot += red[w];
        out = tot;
    }
    __syncthreads();
    return out;
}
// Block-wide selection of the `want` largest keys of keys[0, n) (0 = not a candidate; keys are unique), in
// descending order: sel[r] = index of the r-th largest.  Returns how many were found.
__device__ int block_top_keys(const unsigned long long* keys, int n, int want, int* sel) {
    __shared__ unsigned long long wbest[SAMPLE_THREADS / 32];
    __shared__ unsigned long long s_prev;
    __shared__ int s_found;
    if (threadIdx.x == 0) { s_prev = ~0ull; s_found = 0; }
    __syncthreads();
    for (int r = 0; r < want; ++r) {
        const unsigned long long prev = s_prev;
        unsigned long long best = 0ull;
        for (int c = threadIdx.x; c < n; c += blockDim.x) {
            const unsigned long long k = keys[c];
            if (k < prev && k > best) best = k;
        }
        best = warp_max_u64(best);
        if ((threadIdx.x & 31) == 0) wbest[threadIdx.x >> 5] = best;
        __syncthreads();
        unsigned long long top = wbest[0];
        for (int w = 1; w < (blockDim.x >> 5); ++w) top = wbest[w] > top ? wbest[w] : top;
        if (top == 0ull) break;            // uniform across the block
        for (int c = threadIdx.x; c < n; c += blockDim.x)
            if (keys[c] == top) { sel[r] = c; }
        if (threadIdx.x == 0) { s_prev = top; s_found = r + 1; }
        __syncthreads();
    }
    __syncthreads();
    return s_found;
}

// ---------------------------------------------------------------------------------------------
// sampling select: level + 1 = `width` samples without replacement from q = softmax(flat) where
// flat[j * V + tok] = logp(row j, tok) / T + score(j) over the rows' warped candidates
// ---------------------------------------------------------------------------------------------
__device__ void tree_select_sample_body(const TreeDev& t, const TreeGeom& g, const TrieCSR& trie, int level, int row0,
                                        const int* __restrict__ cand_tok, const int* __restrict__ cand_edge,
                                        const float* __restrict__ cand_logp, const int* __restrict__ cand_cnt, int width,
                                        int P, const SampleCfg& sc, unsigned site) {
    __shared__ unsigned long long keys[SAMPLE_MAX_CAND];
    __shared__ int sel[MAX_BEAMS];
    const int tid = threadIdx.x, B = sc.B, V = g.V;
    auto flat_of = [&](int c) -> float {    // flat[j * V + tok] restricted to the warped candidates; -inf = not a candidate
        const int j = c / B, h = c - j * B;
        if (h >= cand_cnt[row0 + j]) return -INFINITY;
        return cand_logp[(row0 + j) * B + h] * sc.inv_temp + t.score[L_IDX(level, j)];
    };
    const int n_rows = t.cnt[level];
    const int n_cand = n_rows * B;
    const int gen0 = t.scal[SC_GEN0];
    const unsigned long long stream = sc.stream_base | (static_cast<unsigned long long>(level & 0xf) << 4) | site;
    LseAcc acc{-INFINITY, 0.f};
    for (int c = tid; c < n_cand; c += blockDim.x) {
        const float s = flat_of(c);
        if (s > -INFINITY) acc = lse_merge(acc, LseAcc{s, 1.f});
    }
    const float lse = block_lse(acc);
    for (int c = tid; c < n_cand; c += blockDim.x) {
        const float s = flat_of(c);
        unsigned long long k = 0ull;
        if (s > -INFINITY) {
            const int j = c / B, h = c - j * B;
            const uint32_t idx = static_cast<uint32_t>(j) * static_cast<uint32_t>(V) + static_cast<uint32_t>(cand_tok[(row0 + j) * B + h]);
            const float q = expf(s - lse);
            if (q > 0.f) k = rank_key(q / noise_exponential(sc.seed, stream, idx), idx);   // zero-probability entries are never drawn
        }
        keys[c] = k;
    }
    __syncthreads();
    const int n = block_top_keys(keys, n_cand, width, sel);
    const int nl = level + 1;
    for (int p = tid; p < n; p += blockDim.x) {
        const int c = sel[p];
        const int j = c / B, h = c - j * B;
        const int tok = cand_tok[(row0 + j) * B + h], edge = cand_edge[(row0 + j) * B + h];
        const int slot = g.tree_slot(P, nl, p);
        t.tok[L_IDX(nl, p)] = tok;
        t.parent[L_IDX(nl, p)] = j;
        t.score[L_IDX(nl, p)] = flat_of(c);
        t.node[L_IDX(nl, p)] = trie.child_node[edge];
        t.slot[L_IDX(nl, p)] = slot;
        for (int k = 0; k < MAX_NEW; ++k) t.gen[L_IDX(nl, p) * MAX_NEW + k] = t.gen[L_IDX(level, j) * MAX_NEW + k];
        if (gen0 + level < MAX_NEW) t.gen[L_IDX(nl, p) * MAX_NEW + gen0 + level] = tok;
        for (int w = 0; w < VIS_WORDS; ++w) t.vis[L_IDX(nl, p) * VIS_WORDS + w] = t.vis[L_IDX(level, j) * VIS_WORDS + w];
        t.vis[L_IDX(nl, p) * VIS_WORDS + ((slot - P) >> 5)] |= 1u << ((slot - P) & 31);
    }
    if (tid == 0) { t.cnt[nl] = n; t.scal[SC_RESULT] = nl; t.lse_q[level] = lse; }
}

__global__ void __launch_bounds__(SAMPLE_THREADS)
tree_select_sample_kernel(TreeDev t, TreeGeom g, TrieCSR trie, int level, int row0, const int* __restrict__ cand_tok,
                          const int* __restrict__ cand_edge, const float* __restrict__ cand_logp,
                          const int* __restrict__ cand_cnt, int width, int P, SampleCfg sc, unsigned site) {
    tree_select_sample_body(t, g, trie, level, row0, cand_tok, cand_edge, cand_logp, cand_cnt, width, P, sc, site);
}

int tree_select_sample(const TreeDev& t, const TreeGeom& g, const TrieCSR& trie, int level, int row0, const int* cand_tok,
                       const int* cand_edge, const float* cand_logp, const int* cand_cnt, int width, int P,
                       const SampleCfg& sc, unsigned site, cudaStream_t st) {
    ATS_CHECK_ARG(level >= 0 && level + 1 < MAX_LEVELS && width >= 1 && width <= MAX_BEAMS, "select: level=%d width=%d",
                  level, width);
    ATS_CHECK_ARG(sc.B >= 1 && sc.B <= MAX_BEAMS, "select: %d warped candidates per row", sc.B);
    tree_select_sample_kernel<<<1, SAMPLE_THREADS, 0, st>>>(t, g, trie, level, row0, cand_tok, cand_edge, cand_logp,
                                                            cand_cnt, width, P, sc, site);
    ATS_LAUNCH_CHECK();
    return ATS_OK;
}

// ---------------------------------------------------------------------------------------------
// kernel (b), relaxed mode.  Level i: p = softmax of the target's flat scores over the carried beams'
// warped candidates (scattered into the draft's [n_prev, V] index space, beamSD.py:309-321), q = the
// draft's flat softmax of that step; draft pick j is accepted iff r_j <= p_j / q_j (:333-339).
//   >= K accepted: the level is accepted, a random K of them are carried in flat-index order (:340-349);
//   otherwise   : K - #accepted more beams are drawn from norm(max(p - q, 0)) with the accepted
//                 entries removed (:351-369) and the round ends;
//   all levels accepted: the bonus level draws K beams from p (:301-307).
// Where the reference is undefined (empty residual: multinomial over zeros raises for i > 0 and returns
// disallowed tokens for i = 0) the extra beams are drawn from p itself and SC_FALLBACK is incremented.
// ---------------------------------------------------------------------------------------------
__device__ void tree_verify_relaxed_body(const TreeDev& t, const TreeGeom& g, const TrieCSR& trie, int draft_len, int root_rows,
                                         const int* __restrict__ cand_tok, const int* __restrict__ cand_edge,
                                         const float* __restrict__ cand_logp, const int* __restrict__ cand_cnt, int P,
                                         const SampleCfg& sc) {
    __shared__ unsigned long long keys[MAX_K * MAX_BEAMS];
    __shared__ float tflat[MAX_K * MAX_BEAMS];        // target flat score of candidate (k, h)
    __shared__ int sel[MAX_K];
    __shared__ int cur_idx[MAX_K];                    // position of carried beam k in the previous level's list
    __shared__ float cur_score[MAX_K];
    __shared__ int pk_acc[MAX_BEAMS], pk_k[MAX_BEAMS], pk_h[MAX_BEAMS];   // per draft pick: accepted?, carried row, candidate
    __shared__ float pk_s[MAX_BEAMS];
    __shared__ unsigned long long pk_key[MAX_BEAMS];
    __shared__ int chosen[MAX_BEAMS];
    __shared__ int st_idx[MAX_K];
    __shared__ float st_score[MAX_K];
    __shared__ int fin_parent[MAX_K], fin_tok[MAX_K], fin_edge[MAX_K];
    __shared__ float fin_score[MAX_K];
    __shared__ int s_cur_n, s_nacc, s_npk;

    const int tid = threadIdx.x, K = g.K, N = g.N, V = g.V, B = sc.B;
    if (tid == 0) { s_cur_n = t.cnt[0]; s_npk = 0; }
    for (int j = tid; j < t.cnt[0]; j += blockDim.x) { cur_idx[j] = j; cur_score[j] = t.score[L_IDX(0, j)]; }
    __syncthreads();
    int m = 0;
    for (int lvl = 0; lvl <= draft_len; ++lvl) {
        const int rowbase = lvl == 0 ? 0 : root_rows + (lvl - 1) * N;
        const int cur_n = s_cur_n;
        const int n_cand = cur_n * B;
        const unsigned long long lv_stream = sc.stream_base | (static_cast<unsigned long long>(lvl & 0xf) << 4);
        // ---- p: flat target scores of the carried beams' warped candidates ----
        LseAcc acc{-INFINITY, 0.f};
        for (int c = tid; c < n_cand; c += blockDim.x) {
            const int k = c / B, h = c - k * B, row = rowbase + cur_idx[k];
            float s = -INFINITY;
            if (h < cand_cnt[row]) s = cand_logp[row * B + h] * sc.inv_temp + cur_score[k];
            tflat[c] = s;
            if (s > -INFINITY) acc = lse_merge(acc, LseAcc{s, 1.f});
        }
        const float lse_p = block_lse(acc);
        if (lvl == draft_len) {
            // ---- bonus level: K samples from p; index space = [carried rows, V] (beamSD.py:301-307) ----
            for (int c = tid; c < n_cand; c += blockDim.x) {
                unsigned long long key = 0ull;
                if (tflat[c] > -INFINITY) {
                    const int k = c / B, h = c - k * B, row = rowbase + cur_idx[k];
                    const uint32_t idx = static_cast<uint32_t>(k) * V + static_cast<uint32_t>(cand_tok[row * B + h]);
                    const float pb = expf(tflat[c] - lse_p);
                    if (pb > 0.f) key = rank_key(pb / noise_exponential(sc.seed, lv_stream | SITE_BONUS, idx), idx);
                }
                keys[c] = key;
            }
            __syncthreads();
            const int n = block_top_keys(keys, n_cand, K, sel);
            if (tid < n) {
                const int c = sel[tid], k = c / B, h = c - k * B, row = rowbase + cur_idx[k];
                fin_parent[tid] = cur_idx[k]; fin_tok[tid] = cand_tok[row * B + h]; fin_edge[tid] = cand_edge[row * B + h];
                fin_score[tid] = tflat[c];
                t.tr_pick_parent[lvl * MAX_K + tid] = cur_idx[k]; t.tr_pick_tok[lvl * MAX_K + tid] = fin_tok[tid];
                t.tr_pick_score[lvl * MAX_K + tid] = tflat[c]; t.tr_hit_pos[lvl * MAX_K + tid] = -1;
            }
            if (tid == 0) { s_npk = n; t.tr_npick[lvl] = n; }
            __syncthreads();
            break;
        }
        // ---- acceptance test of every draft pick of level lvl + 1 ----
        const int n_next = t.cnt[lvl + 1];
        const float lse_q = t.lse_q[lvl];
        if (tid == 0) s_nacc = 0;
        __syncthreads();
        for (int j = tid; j < n_next; j += blockDim.x) {
            const int pp = t.parent[L_IDX(lvl + 1, j)], tok = t.tok[L_IDX(lvl + 1, j)];
            int k = -1, h = -1;
            for (int kk = 0; kk < cur_n; ++kk) if (cur_idx[kk] == pp) { k = kk; break; }
            float s_t = -INFINITY;
            if (k >= 0) {
                const int row = rowbase + pp, cn = cand_cnt[row];
                for (int hh = 0; hh < cn; ++hh) if (cand_tok[row * B + hh] == tok) { h = hh; break; }
                if (h >= 0) s_t = tflat[k * B + h];
            }
            const float p = s_t > -INFINITY ? expf(s_t - lse_p) : 0.f;
            const float q = expf(t.score[L_IDX(lvl + 1, j)] - lse_q);
            const float r = noise_uniform(sc.seed, lv_stream | SITE_ACCEPT, static_cast<uint32_t>(j));
            const int a = (p > 0.f && r <= p / q) ? 1 : 0;
            pk_acc[j] = a; pk_k[j] = k; pk_h[j] = h; pk_s[j] = s_t;
            pk_key[j] = (static_cast<unsigned long long>(philox_u32(sc.seed, lv_stream | SITE_PERM, static_cast<uint32_t>(j))) << 32) |
                        static_cast<unsigned long long>(j);
            t.tr_acc[L_IDX(lvl, j)] = a;
            if (a) atomicAdd(&s_nacc, 1);
        }
        __syncthreads();
        const int n_acc = s_nacc;
        if (n_acc >= K) {
            // ---- level accepted: a random K of the accepted picks (smallest permutation keys), in flat-index order ----
            for (int j = tid; j < n_next; j += blockDim.x) {
                int c = 0;
                if (pk_acc[j]) {
                    int rank = 0;
                    for (int o = 0; o < n_next; ++o) rank += (pk_acc[o] && pk_key[o] < pk_key[j]) ? 1 : 0;
                    c = rank < K ? 1 : 0;
                }
                chosen[j] = c;
            }
            __syncthreads();
            for (int j = tid; j < n_next; j += blockDim.x) {
                if (!chosen[j]) continue;
                const unsigned int y = static_cast<unsigned int>(t.parent[L_IDX(lvl + 1, j)]) * V + t.tok[L_IDX(lvl + 1, j)];
                int rank = 0;
                for (int o = 0; o < n_next; ++o)
                    if (chosen[o]) {
                        const unsigned int yo = static_cast<unsigned int>(t.parent[L_IDX(lvl + 1, o)]) * V + t.tok[L_IDX(lvl + 1, o)];
                        rank += yo < y ? 1 : 0;
                    }
                st_idx[rank] = j; st_score[rank] = pk_s[j];
                t.tr_pick_parent[lvl * MAX_K + rank] = t.parent[L_IDX(lvl + 1, j)];
                t.tr_pick_tok[lvl * MAX_K + rank] = t.tok[L_IDX(lvl + 1, j)];
                t.tr_pick_score[lvl * MAX_K + rank] = pk_s[j];
                t.tr_hit_pos[lvl * MAX_K + rank] = j;
            }
            __syncthreads();
            for (int r = tid; r < K; r += blockDim.x) { cur_idx[r] = st_idx[r]; cur_score[r] = st_score[r]; }
            if (tid == 0) { s_cur_n = K; t.tr_npick[lvl] = K; }
            ++m;
            __syncthreads();
            continue;
        }
        // ---- level rejected: resample K - n_acc beams from the residual max(p - q, 0) ----
        const int want = K - n_acc;
        float part = 0.f;
        for (int c = tid; c < n_cand; c += blockDim.x) {
            float np = 0.f;
            if (tflat[c] > -INFINITY) {
                const int k = c / B, h = c - k * B, prev = cur_idx[k], row = rowbase + prev;
                const int tok = cand_tok[row * B + h];
                bool accepted = false;
                for (int j = 0; j < n_next; ++j) if (pk_acc[j] && pk_k[j] == k && pk_h[j] == h) { accepted = true; break; }
                if (!accepted) {
                    float q = 0.f;
                    const int dn = t.dcand_cnt[L_IDX(lvl, prev)];
                    // kernel (a) wrote the draft's step-lvl candidates at level base lvl * MAX_BEAMS^2 with row stride B
                    const int* dt = t.dcand_tok + static_cast<long long>(lvl) * MAX_BEAMS * MAX_BEAMS + prev * B;
                    const float* dl = t.dcand_logp + static_cast<long long>(lvl) * MAX_BEAMS * MAX_BEAMS + prev * B;
                    for (int hh = 0; hh < dn; ++hh)
                        if (dt[hh] == tok) { q = expf(dl[hh] * sc.inv_temp + t.score[L_IDX(lvl, prev)] - lse_q); break; }
                    np = fmaxf(expf(tflat[c] - lse_p) - q, 0.f);
                }
            }
            keys[c] = static_cast<unsigned long long>(__float_as_uint(np));   // staged: residual mass as raw bits
            part += np;
        }
        const float tot = block_sum(part);
        const bool fallback = !(tot > 0.f);
        for (int c = tid; c < n_cand; c += blockDim.x) {
            float np = __uint_as_float(static_cast<unsigned int>(keys[c]));
            const int k = c / B, h = c - k * B, prev = cur_idx[k], row = rowbase + prev;
            if (fallback && tflat[c] > -INFINITY) {
                bool accepted = false;
                for (int j = 0; j < n_next; ++j) if (pk_acc[j] && pk_k[j] == k && pk_h[j] == h) { accepted = true; break; }
                np = accepted ? 0.f : expf(tflat[c] - lse_p);
            }
            unsigned long long key = 0ull;
            if (np > 0.f) {
                const uint32_t y = static_cast<uint32_t>(prev) * V + static_cast<uint32_t>(cand_tok[row * B + h]);
                key = rank_key((fallback ? np : np / tot) / noise_exponential(sc.seed, lv_stream | SITE_RESIDUAL, y), y);
            }
            keys[c] = key;     // the same thread staged and now overwrites slot c: no barrier needed
        }
        __syncthreads();
        const int n_extra = block_top_keys(keys, n_cand, want, sel);
        // final beams = accepted picks + extras, ascending flat index (beamSD.py:364-365)
        const int n_fin = n_acc + n_extra;
        for (int j = tid; j < n_next + n_extra; j += blockDim.x) {
            unsigned int y; int par, tok, edge, hit; float sc_t;
            if (j < n_next) {
                if (!pk_acc[j]) continue;
                par = t.parent[L_IDX(lvl + 1, j)]; tok = t.tok[L_IDX(lvl + 1, j)];
                edge = cand_edge[(rowbase + par) * B + pk_h[j]]; sc_t = pk_s[j]; hit = j;
            } else {
                const int c = sel[j - n_next], k = c / B, h = c - k * B;
                par = cur_idx[k]; tok = cand_tok[(rowbase + par) * B + h]; edge = cand_edge[(rowbase + par) * B + h];
                sc_t = tflat[c]; hit = -1;
            }
            y = static_cast<unsigned int>(par) * V + tok;
            int rank = 0;
            for (int o = 0; o < n_next; ++o)
                if (pk_acc[o]) rank += (static_cast<unsigned int>(t.parent[L_IDX(lvl + 1, o)]) * V + t.tok[L_IDX(lvl + 1, o)]) < y ? 1 : 0;
            for (int o = 0; o < n_extra; ++o) {
                const int c = sel[o], k = c / B, h = c - k * B;
                rank += (static_cast<unsigned int>(cur_idx[k]) * V + cand_tok[(rowbase + cur_idx[k]) * B + h]) < y ? 1 : 0;
            }
            fin_parent[rank] = par; fin_tok[rank] = tok; fin_edge[rank] = edge; fin_score[rank] = sc_t;
            t.tr_pick_parent[lvl * MAX_K + rank] = par; t.tr_pick_tok[lvl * MAX_K + rank] = tok;
            t.tr_pick_score[lvl * MAX_K + rank] = sc_t; t.tr_hit_pos[lvl * MAX_K + rank] = hit;
        }
        if (tid == 0) {
            s_npk = n_fin; t.tr_npick[lvl] = n_fin;
            if (fallback) t.scal[SC_FALLBACK] += 1;
        }
        __syncthreads();
        break;
    }
    const int npk = s_npk;
    round_tail(t, g, trie, P, m, npk, draft_len, fin_parent, fin_tok, fin_edge, fin_score);
}

__global__ void __launch_bounds__(SAMPLE_THREADS)
tree_verify_relaxed_kernel(TreeDev t, TreeGeom g, TrieCSR trie, int draft_len, int root_rows,
                           const int* __restrict__ cand_tok, const int* __restrict__ cand_edge,
                           const float* __restrict__ cand_logp, const int* __restrict__ cand_cnt, int P, SampleCfg sc) {
    tree_verify_relaxed_body(t, g, trie, draft_len, root_rows, cand_tok, cand_edge, cand_logp, cand_cnt, P, sc);
}

int tree_verify_relaxed(const TreeDev& t, const TreeGeom& g, const TrieCSR& trie, int draft_len, int root_rows,
                        const int* cand_tok, const int* cand_edge, const float* cand_logp, const int* cand_cnt, int P,
                        const SampleCfg& sc, cudaStream_t st) {
    ATS_CHECK_ARG(draft_len >= 1 && draft_len + 1 < MAX_LEVELS, "verify: draft_len=%d", draft_len);
    ATS_CHECK_ARG(g.K <= MAX_K && sc.B >= 1 && sc.B <= MAX_BEAMS, "verify: K=%d B=%d", g.K, sc.B);
    tree_verify_relaxed_kernel<<<1, SAMPLE_THREADS, 0, st>>>(t, g, trie, draft_len, root_rows, cand_tok, cand_edge,
                                                             cand_logp, cand_cnt, P, sc);
    ATS_LAUNCH_CHECK();
    return ATS_OK;
}

// ---------------------------------------------------------------------------------------------
// final ordering in sampling mode: beams sorted by score, descending (beamSD.py:529-531)
// ---------------------------------------------------------------------------------------------
__device__ void tree_sort_level_body(const TreeDev& t, int level) {
    __shared__ int s_tok[MAX_BEAMS], s_parent[MAX_BEAMS], s_node[MAX_BEAMS], s_slot[MAX_BEAMS], s_gen[MAX_BEAMS][MAX_NEW];
    __shared__ float s_score[MAX_BEAMS];
    __shared__ uint32_t s_vis[MAX_BEAMS][VIS_WORDS];
    const int tid = threadIdx.x, n = t.cnt[level];
    if (tid < n) {
        s_tok[tid] = t.tok[L_IDX(level, tid)]; s_parent[tid] = t.parent[L_IDX(level, tid)];
        s_node[tid] = t.node[L_IDX(level, tid)]; s_slot[tid] = t.slot[L_IDX(level, tid)];
        s_score[tid] = t.score[L_IDX(level, tid)];
        for (int k = 0; k < MAX_NEW; ++k) s_gen[tid][k] = t.gen[L_IDX(level, tid) * MAX_NEW + k];
        for (int w = 0; w < VIS_WORDS; ++w) s_vis[tid][w] = t.vis[L_IDX(level, tid) * VIS_WORDS + w];
    }
    __syncthreads();
    if (tid < n) {
        int rank = 0;
        for (int o = 0; o < n; ++o) rank += (s_score[o] > s_score[tid] || (s_score[o] == s_score[tid] && o < tid)) ? 1 : 0;
        t.tok[L_IDX(level, rank)] = s_tok[tid]; t.parent[L_IDX(level, rank)] = s_parent[tid];
        t.node[L_IDX(level, rank)] = s_node[tid]; t.slot[L_IDX(level, rank)] = s_slot[tid];
        t.score[L_IDX(level, rank)] = s_score[tid];
        for (int k = 0; k < MAX_NEW; ++k) t.gen[L_IDX(level, rank) * MAX_NEW + k] = s_gen[tid][k];
        for (int w = 0; w < VIS_WORDS; ++w) t.vis[L_IDX(level, rank) * VIS_WORDS + w] = s_vis[tid][w];
    }
}

__global__ void __launch_bounds__(64) tree_sort_level_kernel(TreeDev t, int level) { tree_sort_level_body(t, level); }

int tree_sort_level(const TreeDev& t, int level, cudaStream_t st) {
    ATS_CHECK_ARG(level >= 0 && level < MAX_LEVELS, "sort: level=%d", level);
    tree_sort_level_kernel<<<1, 64, 0, st>>>(t, level);
    ATS_LAUNCH_CHECK();
    return ATS_OK;
}

__global__ void noise_fill_kernel(unsigned long long seed, unsigned long long stream, int kind, int n, void* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (kind == 0) static_cast<uint32_t*>(out)[i] = philox_u32(seed, stream, static_cast<uint32_t>(i));
    else if (kind == 1) static_cast<float*>(out)[i] = noise_uniform(seed, stream, static_cast<uint32_t>(i));
    else static_cast<float*>(out)[i] = noise_exponential(seed, stream, static_cast<uint32_t>(i));
}

int noise_fill(unsigned long long seed, unsigned long long stream, int kind, int n, void* out, cudaStream_t st) {
    ATS_CHECK_ARG(n >= 1 && kind >= 0 && kind <= 2 && out, "noise_fill: n=%d kind=%d", n, kind);
    noise_fill_kernel<<<(n + 255) / 256, 256, 0, st>>>(seed, stream, kind, n, out);
    ATS_LAUNCH_CHECK();
    return ATS_OK;
}

}  // namespace atspeed

// =============================================================================================
// Cohort kernels: one launch advances every user of a cohort (grid = users).  Each block runs the
// single-user body above on its user's tree with the user's token / row offsets, so a user's results do
// not depend on who shares the launch with it.
// =============================================================================================
namespace atspeed {

__global__ void cohort_begin_kernel(Cohort c, const TreeDev* __restrict__ trees) {
    const UserCtx& u = c.u[blockIdx.x];
    const TreeDev t = trees[u.tree];
    const int tid = threadIdx.x;
    if (tid < SC_COUNT) t.scal[tid] = 0;
    if (tid < MAX_LEVELS) t.cnt[tid] = tid == 0 ? 1 : 0;
    __syncthreads();
    if (tid == 0) {
        t.scal[SC_P] = u.P;
        t.scal[SC_FIRST] = 1;
        t.tok[0] = -1; t.parent[0] = -1; t.node[0] = 0; t.slot[0] = -1; t.score[0] = 0.f;
    }
    if (tid < MAX_NEW) t.gen[tid] = 0;
    if (tid < VIS_WORDS) t.vis[tid] = 0u;
}

int cohort_begin(const Cohort& c, const TreeDev* trees, cudaStream_t st) {
    ATS_CHECK_ARG(c.n >= 1 && c.n <= MAX_USERS, "cohort of %d users", c.n);
    cohort_begin_kernel<<<c.n, 64, 0, st>>>(c, trees);
    ATS_LAUNCH_CHECK();
    return ATS_OK;
}

__global__ void cohort_build_batch_kernel(Cohort c, const TreeDev* __restrict__ trees, BatchDev b, TreeGeom g,
                                          const int* __restrict__ prompts, int prompt_stride) {
    const UserCtx& u = c.u[blockIdx.x];
    tree_build_batch_body(trees[u.tree], b, g, u.plan, prompts + static_cast<long long>(u.tree) * prompt_stride, u.P, u.T, u.R,
                          u.tok0, u.row0, static_cast<int>(blockIdx.x));
}

// every user's prompt row must start with the n shared-prefix tokens (row `prefix_row` of prompts): *bad |= 1 otherwise
__global__ void cohort_check_prefix_kernel(Cohort c, const int* __restrict__ prompts, int prompt_stride, int prefix_row, int n,
                                           int* __restrict__ bad) {
    const UserCtx& u = c.u[blockIdx.x];
    const int* mine = prompts + static_cast<long long>(u.tree) * prompt_stride;
    const int* pre = prompts + static_cast<long long>(prefix_row) * prompt_stride;
    bool ok = u.P > n;
    for (int i = threadIdx.x; i < n && i < u.P; i += blockDim.x) ok = ok && mine[i] == pre[i];
    if (!ok) atomicOr(bad, 1);
}

int cohort_check_prefix(const Cohort& c, const int* prompts, int prompt_stride, int prefix_row, int n, int* bad, cudaStream_t st) {
    ATS_CHECK_ARG(c.n >= 1 && c.n <= MAX_USERS, "cohort of %d users", c.n);
    cohort_check_prefix_kernel<<<c.n, 64, 0, st>>>(c, prompts, prompt_stride, prefix_row, n, bad);
    ATS_LAUNCH_CHECK();
    return ATS_OK;
}

int cohort_build_batch(const Cohort& c, const TreeDev* trees, const BatchDev& b, const TreeGeom& g, const int* prompts,
                       int prompt_stride, cudaStream_t st) {
    ATS_CHECK_ARG(c.n >= 1 && c.n <= MAX_USERS, "cohort of %d users", c.n);
    cohort_build_batch_kernel<<<c.n, 256, 0, st>>>(c, trees, b, g, prompts, prompt_stride);
    ATS_LAUNCH_CHECK();
    return ATS_OK;
}

__global__ void __launch_bounds__(64)
cohort_select_kernel(Cohort c, const TreeDev* __restrict__ trees, TreeGeom g, TrieCSR trie, int B,
                     const int* __restrict__ cand_tok, const int* __restrict__ cand_edge,
                     const float* __restrict__ cand_logp, const int* __restrict__ cand_cnt) {
    const UserCtx& u = c.u[blockIdx.x];
    if (u.mode != 1) return;
    tree_select_body(trees[u.tree], g, trie, u.level, u.row0, B, cand_tok, cand_edge, cand_logp, cand_cnt, u.width, u.P);
}

__global__ void __launch_bounds__(SAMPLE_THREADS)
cohort_select_sample_kernel(Cohort c, const TreeDev* __restrict__ trees, TreeGeom g, TrieCSR trie,
                            const int* __restrict__ cand_tok, const int* __restrict__ cand_edge,
                            const float* __restrict__ cand_logp, const int* __restrict__ cand_cnt, SampleCfg sc) {
    const UserCtx& u = c.u[blockIdx.x];
    if (u.mode != 1) return;
    const TreeDev t = trees[u.tree];
    sc.stream_base = u.stream_base;
    if (u.is_draft) {
        // keep the draft's warped candidates of this step: the relaxed verify needs q on them (single-user sessions let
        // kernel (a) write there directly; here kernel (a) served the whole cohort into shared rows)
        const int n_rows = t.cnt[u.level];
        const long long lb = static_cast<long long>(u.level) * MAX_BEAMS * MAX_BEAMS;
        for (int i = threadIdx.x; i < n_rows * sc.B; i += blockDim.x) {
            const long long src = static_cast<long long>(u.row0) * sc.B + i;
            t.dcand_tok[lb + i] = cand_tok[src];
            t.dcand_logp[lb + i] = cand_logp[src];
            t.dcand_edge[lb + i] = cand_edge[src];
        }
        for (int i = threadIdx.x; i < n_rows; i += blockDim.x) t.dcand_cnt[u.level * MAX_BEAMS + i] = cand_cnt[u.row0 + i];
    }
    tree_select_sample_body(t, g, trie, u.level, u.row0, cand_tok, cand_edge, cand_logp, cand_cnt, u.width, u.P, sc,
                            u.is_draft ? SITE_DRAFT : SITE_STEP);
}

int cohort_select(const Cohort& c, const TreeDev* trees, const TreeGeom& g, const TrieCSR& trie, int B, const int* cand_tok,
                  const int* cand_edge, const float* cand_logp, const int* cand_cnt, bool sampling, const SampleCfg& sc,
                  cudaStream_t st) {
    ATS_CHECK_ARG(c.n >= 1 && c.n <= MAX_USERS, "cohort of %d users", c.n);
    if (sampling) {
        ATS_CHECK_ARG(sc.B == B, "cohort select: candidate stride %d != sample width %d", B, sc.B);
        cohort_select_sample_kernel<<<c.n, SAMPLE_THREADS, 0, st>>>(c, trees, g, trie, cand_tok, cand_edge, cand_logp, cand_cnt, sc);
    } else {
        cohort_select_kernel<<<c.n, 64, 0, st>>>(c, trees, g, trie, B, cand_tok, cand_edge, cand_logp, cand_cnt);
    }
    ATS_LAUNCH_CHECK();
    return ATS_OK;
}

__global__ void __launch_bounds__(64)
cohort_verify_strict_kernel(Cohort c, const TreeDev* __restrict__ trees, TreeGeom g, TrieCSR trie, int B,
                            const int* __restrict__ cand_tok, const int* __restrict__ cand_edge,
                            const float* __restrict__ cand_logp, const int* __restrict__ cand_cnt) {
    const UserCtx& u = c.u[blockIdx.x];
    if (u.mode != 2) return;
    const long long o = static_cast<long long>(u.row0) * B;
    tree_verify_strict_body(trees[u.tree], g, trie, u.draft_len, u.root_rows, cand_tok + o, cand_edge + o, cand_logp + o,
                            cand_cnt + u.row0, u.P);
}

__global__ void __launch_bounds__(SAMPLE_THREADS)
cohort_verify_relaxed_kernel(Cohort c, const TreeDev* __restrict__ trees, TreeGeom g, TrieCSR trie,
                             const int* __restrict__ cand_tok, const int* __restrict__ cand_edge,
                             const float* __restrict__ cand_logp, const int* __restrict__ cand_cnt, SampleCfg sc) {
    const UserCtx& u = c.u[blockIdx.x];
    if (u.mode != 2) return;
    sc.stream_base = u.stream_base;
    const long long o = static_cast<long long>(u.row0) * sc.B;
    tree_verify_relaxed_body(trees[u.tree], g, trie, u.draft_len, u.root_rows, cand_tok + o, cand_edge + o, cand_logp + o,
                             cand_cnt + u.row0, u.P, sc);
}

int cohort_verify(const Cohort& c, const TreeDev* trees, const TreeGeom& g, const TrieCSR& trie, int B, const int* cand_tok,
                  const int* cand_edge, const float* cand_logp, const int* cand_cnt, bool sampling, const SampleCfg& sc,
                  cudaStream_t st) {
    ATS_CHECK_ARG(c.n >= 1 && c.n <= MAX_USERS, "cohort of %d users", c.n);
    if (sampling)
        cohort_verify_relaxed_kernel<<<c.n, SAMPLE_THREADS, 0, st>>>(c, trees, g, trie, cand_tok, cand_edge, cand_logp, cand_cnt, sc);
    else
        cohort_verify_strict_kernel<<<c.n, 64, 0, st>>>(c, trees, g, trie, B, cand_tok, cand_edge, cand_logp, cand_cnt);
    ATS_LAUNCH_CHECK();
    return ATS_OK;
}

__global__ void cohort_collect_kernel(Cohort c, const TreeDev* __restrict__ trees, int* __restrict__ out4) {
    const int i = threadIdx.x;
    if (i >= c.n) return;
    const TreeDev& t = trees[c.u[i].tree];
    out4[4 * i + 0] = t.scal[SC_NMATCH];
    out4[4 * i + 1] = t.scal[SC_MISS];
    out4[4 * i + 2] = t.cnt[0];
    out4[4 * i + 3] = t.scal[SC_FALLBACK];
}

int cohort_collect(const Cohort& c, const TreeDev* trees, int* out4, cudaStream_t st) {
    cohort_collect_kernel<<<1, 32, 0, st>>>(c, trees, out4);
    ATS_LAUNCH_CHECK();
    return ATS_OK;
}

// u.level = the level that holds the user's final beams, u.row0 = the record index the user's result is written to
__global__ void __launch_bounds__(64)
cohort_results_kernel(Cohort c, const TreeDev* __restrict__ trees, int K, int sort, int* __restrict__ tokens,
                      float* __restrict__ scores, int* __restrict__ counts) {
    const UserCtx& u = c.u[blockIdx.x];
    const TreeDev t = trees[u.tree];
    const long long o = u.row0;
    if (sort) {
        tree_sort_level_body(t, u.level);
        __syncthreads();
    }
    const int n = t.cnt[u.level] < K ? t.cnt[u.level] : K;
    for (int i = threadIdx.x; i < K * MAX_NEW; i += blockDim.x) {
        const int b = i / MAX_NEW, k = i - b * MAX_NEW;
        tokens[(o * K + b) * MAX_NEW + k] = b < n ? t.gen[L_IDX(u.level, b) * MAX_NEW + k] : 0;
    }
    for (int b = threadIdx.x; b < K; b += blockDim.x) scores[o * K + b] = b < n ? t.score[L_IDX(u.level, b)] : -INFINITY;
    if (threadIdx.x == 0 && counts) counts[o] = n;
}

int cohort_results(const Cohort& c, const TreeDev* trees, int K, bool sort, int* tokens, float* scores, int* counts,
                   cudaStream_t st) {
    ATS_CHECK_ARG(c.n >= 1 && c.n <= MAX_USERS, "cohort of %d users", c.n);
    cohort_results_kernel<<<c.n, 64, 0, st>>>(c, trees, K, sort ? 1 : 0, tokens, scores, counts);
    ATS_LAUNCH_CHECK();
    return ATS_OK;
}

}  // namespace atspeed
