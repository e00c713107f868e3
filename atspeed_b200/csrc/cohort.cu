// Cohort scheduler: speculative beam search for MANY users with shared forwards ("multi-user batching of
// the tree forward"; the reference is batch-1 only: code/utils.py:103, code/beamSD.py:57,224).
//
// A single search streams the 13.5 GB of target weights once per round for <= 220 tokens -- a B200 is idle on every axis
// but HBM while it does so.  Users are independent, so the scheduler below keeps up to cfg.max_users searches in flight
// and, at every step, packs the draft steps / verify forwards / final steps of all users that are ready into ONE forward
// of at most 512 tokens per model (block-diagonal tree masks: every token only sees its own user's KV cache), followed
// by ONE launch each of kernel (a), the select / verify kernels (grid = users) and kernel (c).  Per-user arithmetic is
// exactly that of the single-user session (same kernels' bodies, same candidate lists), so ranked lists, accepted
// lengths and scores do not depend on who shares a forward -- tests/test_gpu_cohort.py checks that against single-user
// runs.  One stream synchronisation per scheduler step returns the accepted lengths of the users verified in it.
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "session.h"

using namespace atspeed;

namespace {

struct UserRun {
    int req;            // index in the caller's arrays
    int slot;           // user slot: tree, KV caches, prompt row
    int P;
    int done;           // tokens generated so far
    bool first, miss;   // roots are still the prompt / the draft owes KV for accepted tokens
    int dl;             // draft length of the current round; 0 = final plain target step (code/beamSD.py:505-509)
    int j;              // draft steps done in this round
    bool finished;
    int level;          // tree level that holds the final beams
    int round;
    unsigned long long user_seq;
    int n_run, total, acc[8], tf, df;
    int waited;         // scheduler steps this user's target forward has been held back for a fuller pack
};

struct Pack {
    Cohort c;
    std::vector<int> who;     // indices into the active list
    int T, R;
};

// token / row extent and batch plan of one draft step (mirrors search_step in engine.cu)
void plan_draft_step(const atspeed_session* s, const UserRun& u, UserCtx& x, int& S) {
    const TreeGeom& g = s->geom;
    memset(&x.plan, 0, sizeof(x.plan));
    x.level = u.j; x.width = g.N; x.mode = 1; x.is_draft = 1; x.draft_len = 0; x.root_rows = 0;
    if (u.j == 0 && u.first) {
        x.plan.with_prompt = 1; x.plan.l_from = 1; x.plan.l_to = 0; x.plan.rows_from = 1; x.plan.root_row = 1; x.plan.width = g.N;
        x.plan.prompt_skip = s->prefix_len;
        x.T = u.P - s->prefix_len; x.R = 1; S = u.P;
    } else {
        const bool with_missing = u.j == 0 && u.miss;
        x.plan.with_missing = with_missing ? 1 : 0;
        x.plan.l_from = x.plan.l_to = x.plan.rows_from = u.j; x.plan.width = g.N;
        const int cap = u.j == 0 ? g.K : g.N;
        x.T = (with_missing ? g.K : 0) + cap; x.R = cap;
        S = g.tree_slot(u.P, u.j, 0) + cap;
    }
}

// the target's verify forward (atspeed_session_target) or, with dl == 0, its final plain step
void plan_target(const atspeed_session* s, const UserRun& u, UserCtx& x, int& S) {
    const TreeGeom& g = s->geom;
    memset(&x.plan, 0, sizeof(x.plan));
    x.is_draft = 0;
    if (u.dl == 0) {
        x.level = 0; x.width = g.K; x.mode = 1; x.draft_len = 0; x.root_rows = 0;
        if (u.first) {
            x.plan.with_prompt = 1; x.plan.l_from = 1; x.plan.l_to = 0; x.plan.rows_from = 1; x.plan.root_row = 1; x.plan.width = g.K;
            x.plan.prompt_skip = s->prefix_len;
            x.T = u.P - s->prefix_len; x.R = 1; S = u.P;
        } else {
            x.plan.l_from = x.plan.l_to = x.plan.rows_from = 0; x.plan.width = g.K;
            x.T = g.K; x.R = g.K; S = g.tree_slot(u.P, 0, 0) + g.K;
        }
        return;
    }
    x.level = 0; x.width = g.N; x.mode = 2; x.draft_len = u.dl; x.root_rows = u.first ? 1 : g.K;
    x.plan.width = g.N;
    if (u.first) {
        x.plan.with_prompt = 1; x.plan.l_from = 1; x.plan.l_to = u.dl; x.plan.rows_from = 1; x.plan.root_row = 1;
        x.plan.prompt_skip = s->prefix_len;
        x.T = u.P - s->prefix_len + u.dl * g.N; x.R = 1 + u.dl * g.N;
    } else {
        x.plan.l_from = 0; x.plan.l_to = u.dl; x.plan.rows_from = 0;
        x.T = g.K + u.dl * g.N; x.R = x.T;
    }
    S = g.tree_slot(u.P, u.dl, 0) + g.N;
}

}  // namespace

namespace atspeed {

// Packing of the target forwards of one scheduler step (pure host arithmetic; atspeed_debug_plan_packs exposes it to the CPU
// tests).  A target forward's cost is dominated by how many forwards carry the tokens, so ready users are packed best-fit
// decreasing: items sorted by token count (stable), every pack opened by the largest item left and filled with the largest
// that still fit (tokens <= T_max, logit rows <= R_max, users <= max_users).  A pack that fills less than 7/8 of the forward
// is held back one step -- the users launched now come back with small later-round trees (K + dl N tokens) that fill it --
// unless deferral is off, no new user can arrive (no_more_work), one of its users has already waited twice, or nothing at
// all would run (then the fullest pack runs).  order[k] = index of the k-th item in packing order; pack_of[i] = pack of item
// i (packs numbered in the order they are opened); run_now[b] = 1 when pack b runs in this step.  Returns the pack count.
int plan_packs(const PackItem* items, int n, int T_max, int R_max, int max_users, bool defer, bool no_more_work, int* order,
               int* pack_of, unsigned char* run_now) {
    for (int i = 0; i < n; ++i) { order[i] = i; pack_of[i] = -1; }
    std::stable_sort(order, order + n, [&](int p, int q) { return items[p].T > items[q].T; });
    int n_packs = 0;
    std::vector<int> pack_T;
    for (int first = 0; first < n; ++first) {
        if (pack_of[order[first]] >= 0) continue;
        int T = 0, R = 0, users = 0;
        for (int k = first; k < n; ++k) {
            const int i = order[k];
            if (pack_of[i] >= 0) continue;
            if (users == max_users || T + items[i].T > T_max || R + items[i].R > R_max) continue;
            pack_of[i] = n_packs; T += items[i].T; R += items[i].R; ++users;
        }
        if (users == 0) { pack_of[order[first]] = n_packs; T = items[order[first]].T; }   // an item larger than a forward: the caller rejects it
        pack_T.push_back(T);
        ++n_packs;
    }
    const int full = T_max - T_max / 8;
    int launched = 0, fullest = -1;
    for (int b = 0; b < n_packs; ++b) {
        bool aged = false;
        for (int i = 0; i < n; ++i) aged = aged || (pack_of[i] == b && items[i].waited >= 2);
        run_now[b] = (!defer || pack_T[b] >= full || no_more_work || aged) ? 1 : 0;
        launched += run_now[b];
        if (fullest < 0 || pack_T[b] > pack_T[fullest]) fullest = b;
    }
    if (launched == 0 && fullest >= 0) run_now[fullest] = 1;          // progress
    return n_packs;
}

// run one packed forward of model `m` + kernel (a) + the per-user select / verify kernels
static int run_pack(atspeed_session* s, ModelRT& m, const Pack& pk, const CohortKV& ckv, int B, bool has_verify,
                    bool has_select, cudaStream_t st) {
    const bool sampling = s->cfg.do_sample != 0;
    PROF(s, CAT_BEAM, 0, cohort_build_batch(pk.c, s->trees_dev, s->batch, s->geom, s->prompts_dev, s->cfg.max_prompt, st));
    s->launches += 1;
    BatchDesc b;
    memset(&b, 0, sizeof(b));
    b.tok = s->batch.tok; b.pos = s->batch.pos; b.slot = s->batch.slot; b.prefix_len = s->batch.prefix_len;
    b.vis = s->batch.vis; b.vis_base = 0; b.n_valid = nullptr; b.tok_user = s->batch.tok_user; b.ckv = ckv;
    int max_S = 1;
    for (int i = 0; i < ckv.n; ++i) max_S = ckv.S[i] > max_S ? ckv.S[i] : max_S;
    ATS_TRY(forward(s, m, b, pk.T, max_S, s->batch.rows_idx, pk.R, st));
    ATS_TRY(run_topk(s, m, pk.R, B, shared_cand(s), st));
    SampleCfg sc = sample_cfg(s);
    if (has_select) {
        PROF(s, CAT_BEAM, 0,
             cohort_select(pk.c, s->trees_dev, s->geom, s->trie, B, s->cand_tok, s->cand_edge, s->cand_logp, s->cand_cnt,
                           sampling, sc, st));
        s->launches += 1;
    }
    if (has_verify) {
        PROF(s, CAT_BEAM, 0,
             cohort_verify(pk.c, s->trees_dev, s->geom, s->trie, B, s->cand_tok, s->cand_edge, s->cand_logp, s->cand_cnt,
                           sampling, sc, st));
        s->launches += 1;
    }
    return ATS_OK;
}

}  // namespace atspeed

// device_io: prompts (concatenated) already live in HBM and the final beams stay there (tokens [n][K][MAX_NEW] int32,
// scores [n][K] fp32, record i = request i); otherwise both are host buffers (tokens [n][K * max_new_tokens]).
static int bssd_batch_impl(atspeed_session* s, int32_t n_users, const int32_t* prompts, const int32_t* prompt_lens,
                           int32_t gamma, int32_t* tokens_host, float* scores_host, int32_t* counts, int32_t* tokens_dev,
                           float* scores_dev, atspeed_stats* stats, bool device_io, void* stream) {
    const int32_t* prompts_host = prompts;
    ATS_CHECK_ARG(s && s->has_draft, "session has no draft model");
    ATS_CHECK_ARG(s->max_users > 1, "atspeed_bssd_batch needs a session created with max_users > 1");
    ATS_CHECK_ARG(n_users >= 1 && prompts && prompt_lens, "null / empty request");
    ATS_CHECK_ARG(device_io ? (tokens_dev && scores_dev) : (tokens_host && scores_host && counts), "null result buffers");
    ATS_CHECK_ARG(gamma >= 1, "gamma=%d", gamma);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const TreeGeom& g = s->geom;
    const int L = s->cfg.max_new_tokens, K = g.K, U = s->max_users;
    const bool sampling = s->cfg.do_sample != 0;
    std::vector<long long> poff(n_users + 1, 0);
    for (int i = 0; i < n_users; ++i) {
        const int P = prompt_lens[i];
        ATS_CHECK_ARG(P >= 1 && P <= s->cfg.max_prompt, "prompt %d: length %d outside [1,%d]", i, P, s->cfg.max_prompt);
        ATS_CHECK_ARG(P + (L - 1) * g.N <= s->T_max && P + K <= s->T_max, "prompt %d: %d + %d tree tokens exceed the %d-token forward",
                      i, P, (L - 1) * g.N, s->T_max);
        ATS_CHECK_ARG(P > s->prefix_len, "prompt %d: length %d does not exceed the shared prefix (%d tokens)", i, P, s->prefix_len);
        poff[i + 1] = poff[i] + P;
    }
    if (s->prefix_len > 0) ATS_CUDA(cudaMemsetAsync(s->prefix_bad_dev, 0, sizeof(int), static_cast<cudaStream_t>(stream)));
    // pinned staging (owned by the session) for the per-step outcomes and the final beams of one cohort
    ATS_CHECK_ARG(s->cohort_pinned, "session has no cohort staging buffer");
    int* h_collect = s->cohort_pinned;
    int* h_tok = h_collect + MAX_USERS * 4;
    int* h_cnt = h_tok + MAX_USERS * MAX_K * MAX_NEW;
    float* h_score = reinterpret_cast<float*>(h_cnt + MAX_USERS);
    int* h_prefix_bad = reinterpret_cast<int*>(h_score + MAX_USERS * MAX_K);
    *h_prefix_bad = 0;

    static const bool log_packs = []() { const char* e = getenv("ATSPEED_COHORT_LOG"); return e && atoi(e) == 1; }();   // diagnostics
    static const bool pack_defer = []() { const char* e = getenv("ATSPEED_COHORT_DEFER"); return !(e && atoi(e) == 0); }();   // A/B: 0 = run every pack at once
    std::vector<UserRun> act;
    std::vector<int> free_slots;
    for (int u = U - 1; u >= 0; --u) free_slots.push_back(u);
    int next = 0, finished = 0;
    const long long l0 = s->launches;

    auto start_round = [&](UserRun& u) {
        int dl = gamma < L - u.done - 1 ? gamma : L - u.done - 1;          // code/beamSD.py:504
        if (dl > MAX_LEVELS - 1) dl = MAX_LEVELS - 1;
        u.dl = dl; u.j = 0;
    };
    auto ckv_entry = [&](CohortKV& ckv, int i, const UserRun& u, const UserCtx& x, int S, long long kv_user_elems) {
        ckv.kv_off[i] = static_cast<long long>(u.slot) * kv_user_elems;
        ckv.S[i] = S; ckv.vis_base[i] = u.P; ckv.tok0[i] = x.tok0; ckv.T[i] = x.T;
    };

    while (finished < n_users) {
        // ---- admit new users into free slots ----
        {
            Cohort c;
            memset(&c, 0, sizeof(c));
            while (!free_slots.empty() && next < n_users) {
                UserRun u;
                memset(&u, 0, sizeof(u));
                u.req = next; u.slot = free_slots.back(); free_slots.pop_back();
                u.P = prompt_lens[next]; u.first = true; u.user_seq = s->next_user_seq++;
                ATS_CUDA(cudaMemcpyAsync(s->prompts_dev + static_cast<long long>(u.slot) * s->cfg.max_prompt,
                                         prompts_host + poff[next], sizeof(int) * u.P,
                                         device_io ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
                start_round(u);
                c.u[c.n].tree = u.slot; c.u[c.n].P = u.P;
                ++c.n;
                act.push_back(u);
                ++next;
                if (s->prefix_len > 0) {
                    // the shared prefix's K/V rows -> slots [0, prefix_len) of this user's caches (kernel (c), out of place)
                    for (ModelRT* m : {&s->tgt, &s->dft}) {
                        const long long per_user = (m == &s->tgt ? s->kv_user_elems_tgt : s->kv_user_elems_dft) * m->elem_bytes;
                        uint8_t* base = m->f32 ? reinterpret_cast<uint8_t*>(m->fkv) : reinterpret_cast<uint8_t*>(m->kv);
                        PROF(s, CAT_GATHER, 0,
                             kv_gather_rows_oop(base + static_cast<long long>(U) * per_user, base + static_cast<long long>(u.slot) * per_user,
                                                m->kv_plane * m->elem_bytes, m->kv_plane * m->elem_bytes, m->d.n_layers * 2,
                                                m->HD * m->elem_bytes, s->prefix_iota_dev, s->prefix_iota_dev, nullptr, s->prefix_len, st));
                        s->launches += 1;
                    }
                }
            }
            if (c.n > 0) {
                PROF(s, CAT_BEAM, 0, cohort_begin(c, s->trees_dev, st));
                s->launches += 1;
                if (s->prefix_len > 0) {
                    PROF(s, CAT_BEAM, 0, cohort_check_prefix(c, s->prompts_dev, s->cfg.max_prompt, U, s->prefix_len, s->prefix_bad_dev, st));
                    s->launches += 1;
                }
            }
        }
        // ---- draft: every user with draft steps left runs its next step; repeat until none has ----
        for (;;) {
            Pack pk;
            memset(&pk.c, 0, sizeof(pk.c));
            pk.T = pk.R = 0;
            CohortKV ckv;
            memset(&ckv, 0, sizeof(ckv));
            for (size_t a = 0; a < act.size(); ++a) {
                UserRun& u = act[a];
                if (u.finished || u.j >= u.dl) continue;
                UserCtx x;
                memset(&x, 0, sizeof(x));
                int S = 0;
                plan_draft_step(s, u, x, S);
                if (pk.c.n == MAX_USERS || pk.T + x.T > s->T_max || pk.R + x.R > s->R_max) continue;   // next forward
                x.tree = u.slot; x.P = u.P; x.tok0 = pk.T; x.row0 = pk.R;
                x.stream_base = noise_stream(u.user_seq, static_cast<uint32_t>(u.round), 0, 0);
                ckv_entry(ckv, pk.c.n, u, x, S, s->kv_user_elems_dft);
                pk.c.u[pk.c.n++] = x;
                pk.T += x.T; pk.R += x.R;
                pk.who.push_back(static_cast<int>(a));
            }
            if (pk.c.n == 0) break;
            ckv.n = pk.c.n;
            ATS_TRY(run_pack(s, s->dft, pk, ckv, sampling ? s->sample_B : g.N, false, true, st));
            for (int a : pk.who) { act[a].j += 1; act[a].df += 1; }
        }
        // ---- target: verify forwards (dl >= 1) and final steps (dl == 0) of the users that are ready ----
        // Packing: plan_packs (above) decides who shares a forward and which packs wait a step for a better fill.  Who shares a
        // forward never changes a user's results.
        bool any_target = false;
        std::vector<Pack> packs;
        std::vector<CohortKV> pack_kv;
        std::vector<int> pack_flags;       // bit 0: has_verify, bit 1: has_select; max_dl in bits 8..
        std::vector<unsigned char> run_now;
        {
            std::vector<int> who;          // index into act of every ready user
            std::vector<PackItem> items;
            for (size_t a = 0; a < act.size(); ++a) {
                UserRun& u = act[a];
                if (u.finished || u.j < u.dl) continue;
                UserCtx x;
                memset(&x, 0, sizeof(x));
                int S = 0;
                plan_target(s, u, x, S);
                who.push_back(static_cast<int>(a));
                items.push_back(PackItem{x.T, x.R, u.waited});
            }
            const int n = static_cast<int>(items.size());
            std::vector<int> order(n), pack_of(n);
            run_now.assign(n > 0 ? n : 1, 0);
            const int n_packs = plan_packs(items.data(), n, s->T_max, s->R_max, MAX_USERS, pack_defer, next >= n_users, order.data(),
                                           pack_of.data(), run_now.data());
            packs.resize(n_packs);
            pack_kv.resize(n_packs);
            pack_flags.assign(n_packs, 0);
            for (int b = 0; b < n_packs; ++b) {
                memset(&packs[b].c, 0, sizeof(packs[b].c));
                packs[b].T = packs[b].R = 0;
                memset(&pack_kv[b], 0, sizeof(CohortKV));
            }
            for (int k = 0; k < n; ++k) {
                const int i = order[k], b = pack_of[i];
                Pack& pk = packs[b];
                UserRun& u = act[who[i]];
                UserCtx x;
                memset(&x, 0, sizeof(x));
                int S = 0;
                plan_target(s, u, x, S);
                x.tree = u.slot; x.P = u.P; x.tok0 = pk.T; x.row0 = pk.R;
                x.stream_base = noise_stream(u.user_seq, static_cast<uint32_t>(u.round), 0, 0);
                ckv_entry(pack_kv[b], pk.c.n, u, x, S, s->kv_user_elems_tgt);
                pk.c.u[pk.c.n++] = x;
                pk.T += x.T; pk.R += x.R;
                pk.who.push_back(who[i]);
                int max_dl = pack_flags[b] >> 8, flags = pack_flags[b] & 0xff;
                if (x.mode == 2) { flags |= 1; max_dl = u.dl > max_dl ? u.dl : max_dl; } else flags |= 2;
                pack_flags[b] = flags | (max_dl << 8);
                pack_kv[b].n = pk.c.n;
                u.waited = run_now[b] ? 0 : u.waited + 1;
            }
        }
        for (size_t pb = 0; pb < packs.size(); ++pb) {
            if (!run_now[pb]) continue;
            Pack& pk = packs[pb];
            CohortKV& ckv = pack_kv[pb];
            const bool has_verify = (pack_flags[pb] & 1) != 0, has_select = (pack_flags[pb] & 2) != 0;
            const int max_dl = pack_flags[pb] >> 8;
            if (pk.c.n == 0) continue;
            any_target = true;
            ckv.n = pk.c.n;
            if (log_packs) fprintf(stderr, "atspeed-pack target T=%d R=%d users=%d active=%zu\n", pk.T, pk.R, pk.c.n, act.size());
            ATS_TRY(run_pack(s, s->tgt, pk, ckv, sampling ? s->sample_B : K, has_verify, has_select, st));
            // kernel (c): move the survivors' ancestor rows into the accepted region, both caches, verified users only
            if (has_verify) {
                for (ModelRT* m : {&s->tgt, &s->dft}) {
                    GatherCohort gc;
                    memset(&gc, 0, sizeof(gc));
                    const long long per_user = m == &s->tgt ? s->kv_user_elems_tgt : s->kv_user_elems_dft;
                    for (int i = 0; i < pk.c.n; ++i) {
                        if (pk.c.u[i].mode != 2) continue;
                        const TreeDev& t = s->trees_host[pk.c.u[i].tree];
                        gc.byte_off[gc.n] = static_cast<long long>(pk.c.u[i].tree) * per_user * m->elem_bytes;
                        gc.src[gc.n] = t.gather_src; gc.dst[gc.n] = t.gather_dst; gc.n_rows[gc.n] = t.scal + SC_GATHER;
                        ++gc.n;
                    }
                    PROF(s, CAT_GATHER, 0,
                         kv_gather_rows_cohort(m->f32 ? static_cast<void*>(m->fkv) : static_cast<void*>(m->kv), m->kv_plane * m->elem_bytes, m->d.n_layers * 2, m->HD * m->elem_bytes, gc,
                                               (max_dl + 1) * K, st));
                    s->launches += 1;
                }
            }
            PROF(s, CAT_BEAM, 0, cohort_collect(pk.c, s->trees_dev, s->collect_dev, st));
            s->launches += 1;
            ATS_CUDA(cudaMemcpyAsync(h_collect, s->collect_dev, sizeof(int) * 4 * pk.c.n, cudaMemcpyDeviceToHost, st));
            if (s->prefix_len > 0) ATS_CUDA(cudaMemcpyAsync(h_prefix_bad, s->prefix_bad_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
            // users that finish in this pack: results straight away (their slots are refilled next step)
            Cohort fin;
            memset(&fin, 0, sizeof(fin));
            std::vector<int> fin_who;
            ATS_CUDA(cudaStreamSynchronize(st));
            ATS_CHECK_ARG(s->prefix_len == 0 || *h_prefix_bad == 0,
                          "a prompt does not start with the %d tokens given to atspeed_session_set_shared_prefix", s->prefix_len);
            for (int i = 0; i < pk.c.n; ++i) {
                UserRun& u = act[pk.who[i]];
                u.tf += 1;
                if (pk.c.u[i].mode == 2) {
                    const int mch = h_collect[4 * i + 0];
                    if (u.n_run < 8) u.acc[u.n_run] = mch;
                    u.n_run += 1; u.total += mch; u.done += mch + 1;
                    u.first = false; u.miss = h_collect[4 * i + 1] > 0; u.round += 1; u.level = 0;
                    if (u.done >= L) u.finished = true; else start_round(u);
                } else {
                    u.finished = true; u.level = 1;             // the final step's beams are level 1
                }
                if (u.finished) {
                    fin.u[fin.n].tree = u.slot; fin.u[fin.n].level = u.level; fin.u[fin.n].P = u.P;
                    fin.u[fin.n].row0 = device_io ? u.req : fin.n;       // record index of the result
                    ++fin.n;
                    fin_who.push_back(pk.who[i]);
                }
            }
            if (fin.n > 0) {
                if (device_io) {
                    PROF(s, CAT_BEAM, 0, cohort_results(fin, s->trees_dev, K, sampling, tokens_dev, scores_dev, nullptr, st));
                } else {
                    PROF(s, CAT_BEAM, 0, cohort_results(fin, s->trees_dev, K, sampling, s->res_tok_dev, s->res_score_dev, s->res_cnt_dev, st));
                    ATS_CUDA(cudaMemcpyAsync(h_tok, s->res_tok_dev, sizeof(int) * fin.n * K * MAX_NEW, cudaMemcpyDeviceToHost, st));
                    ATS_CUDA(cudaMemcpyAsync(h_score, s->res_score_dev, sizeof(float) * fin.n * K, cudaMemcpyDeviceToHost, st));
                    ATS_CUDA(cudaMemcpyAsync(h_cnt, s->res_cnt_dev, sizeof(int) * fin.n, cudaMemcpyDeviceToHost, st));
                    ATS_CUDA(cudaStreamSynchronize(st));
                }
                s->launches += 1;
                for (int i = 0; i < fin.n; ++i) {
                    UserRun& u = act[fin_who[i]];
                    if (!device_io) {
                        const int n = h_cnt[i];
                        counts[u.req] = n;
                        for (int b = 0; b < K; ++b) {
                            for (int k = 0; k < L; ++k)
                                tokens_host[(static_cast<long long>(u.req) * K + b) * L + k] = h_tok[(i * K + b) * MAX_NEW + k];
                            scores_host[static_cast<long long>(u.req) * K + b] = h_score[i * K + b];
                        }
                    }
                    if (stats) {
                        atspeed_stats& o = stats[u.req];
                        memset(&o, 0, sizeof(o));
                        o.n_run = u.n_run; o.total_accept_steps = u.total;
                        for (int r = 0; r < 8; ++r) o.accept_steps[r] = u.acc[r];
                        o.target_forwards = u.tf; o.draft_forwards = u.df;
                    }
                    free_slots.push_back(u.slot);
                    ++finished;
                }
            }
        }
        // drop finished users from the active list
        act.erase(std::remove_if(act.begin(), act.end(), [](const UserRun& u) { return u.finished; }), act.end());
        if (!any_target && act.empty() && next >= n_users && finished < n_users) {
            set_error("cohort scheduler stalled with %d of %d users finished", finished, n_users);
            return ATS_ERR_STATE;
        }
    }
    if (stats) {
        // kernel launches are shared by the cohort: report the per-user average
        const int per_user = static_cast<int>((s->launches - l0) / n_users);
        for (int i = 0; i < n_users; ++i) stats[i].kernel_launches = per_user;
    }
    return ATS_OK;
}

// CPU-testable view of the scheduler's packing policy (tests/test_cohort_packing.py)
extern "C" int atspeed_debug_plan_packs(const int32_t* T, const int32_t* R, const int32_t* waited, int32_t n, int32_t T_max,
                                        int32_t R_max, int32_t defer, int32_t no_more_work, int32_t* pack_of, uint8_t* run_now,
                                        int32_t* n_packs) {
    ATS_CHECK_ARG(n >= 0 && n <= 4096 && (n == 0 || (T && R && waited && pack_of)) && run_now && n_packs && T_max >= 1 && R_max >= 1,
                  "plan_packs: bad arguments");
    std::vector<PackItem> items(n);
    for (int i = 0; i < n; ++i) items[i] = PackItem{T[i], R[i], waited[i]};
    std::vector<int> order(n);
    *n_packs = plan_packs(items.data(), n, T_max, R_max, MAX_USERS, defer != 0, no_more_work != 0, order.data(), pack_of, run_now);
    return ATS_OK;
}

// Shared prompt prefix: every prompt of the reference's datasets opens with the same instruction template (39 tokens of ~105 on
// Beauty / Games).  Its K/V rows do not depend on the user, so they are computed ONCE per session and model here -- one forward
// of the prefix in the extra user slot -- and copied into a user's caches on admission; the user's own forwards then start at
// token `n`.  Exact in the fp32 parity mode; in bf16 the prefix rows come from a forward of a different size, which may flip
// the same near-ties any other batch-composition change does.  n = 0 switches it off.
extern "C" int atspeed_session_set_shared_prefix(atspeed_session* s, const int32_t* prefix_host, int32_t n, void* stream) {
    ATS_CHECK_ARG(s && s->max_users > 1, "a shared prefix needs a cohort session (max_users > 1)");
    ATS_CHECK_ARG(n >= 0 && n < s->cfg.max_prompt && n <= s->T_max && (n == 0 || prefix_host), "shared prefix of %d tokens", n);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    s->prefix_len = 0;
    if (n == 0) return ATS_OK;
    const int U = s->max_users;
    {
        std::vector<int> iota(s->cfg.max_prompt);
        for (int i = 0; i < s->cfg.max_prompt; ++i) iota[i] = i;
        ATS_CUDA(cudaMemcpyAsync(s->prefix_iota_dev, iota.data(), sizeof(int) * iota.size(), cudaMemcpyHostToDevice, st));
        ATS_CUDA(cudaMemcpyAsync(s->prompts_dev + static_cast<long long>(U) * s->cfg.max_prompt, prefix_host, sizeof(int) * n,
                                 cudaMemcpyHostToDevice, st));
        ATS_CUDA(cudaStreamSynchronize(st));          // the host vectors go out of scope
    }
    Cohort c;
    memset(&c, 0, sizeof(c));
    c.n = 1; c.u[0].tree = U; c.u[0].P = n;
    ATS_TRY(cohort_begin(c, s->trees_dev, st));
    for (ModelRT* m : {&s->tgt, &s->dft}) {
        if (m == &s->dft && !s->has_draft) continue;
        Pack pk;
        memset(&pk.c, 0, sizeof(pk.c));
        UserCtx& x = pk.c.u[0];
        x.plan.with_prompt = 1; x.plan.l_from = 1; x.plan.l_to = 0; x.plan.rows_from = 1; x.plan.root_row = 1; x.plan.width = s->geom.K;
        x.level = 0; x.width = s->geom.K; x.mode = 0;
        x.tree = U; x.P = n; x.tok0 = 0; x.row0 = 0; x.T = n; x.R = 1;
        pk.c.n = 1; pk.T = n; pk.R = 1;
        CohortKV ckv;
        memset(&ckv, 0, sizeof(ckv));
        ckv.n = 1;
        ckv.kv_off[0] = static_cast<long long>(U) * (m == &s->tgt ? s->kv_user_elems_tgt : s->kv_user_elems_dft);
        ckv.S[0] = n; ckv.vis_base[0] = n; ckv.tok0[0] = 0; ckv.T[0] = n;
        ATS_TRY(cohort_build_batch(pk.c, s->trees_dev, s->batch, s->geom, s->prompts_dev, s->cfg.max_prompt, st));
        BatchDesc b;
        memset(&b, 0, sizeof(b));
        b.tok = s->batch.tok; b.pos = s->batch.pos; b.slot = s->batch.slot; b.prefix_len = s->batch.prefix_len;
        b.vis = s->batch.vis; b.vis_base = 0; b.n_valid = nullptr; b.tok_user = s->batch.tok_user; b.ckv = ckv;
        ATS_TRY(forward(s, *m, b, n, n, s->batch.rows_idx, 1, st));
    }
    ATS_CUDA(cudaStreamSynchronize(st));
    s->prefix_len = n;
    return ATS_OK;
}

extern "C" int atspeed_bssd_batch(atspeed_session* s, int32_t n_users, const int32_t* prompts_host,
                                  const int32_t* prompt_lens, int32_t gamma, int32_t* tokens_host, float* scores_host,
                                  int32_t* counts, atspeed_stats* stats, void* stream) {
    return bssd_batch_impl(s, n_users, prompts_host, prompt_lens, gamma, tokens_host, scores_host, counts, nullptr, nullptr,
                           stats, false, stream);
}

extern "C" int atspeed_bssd_batch_device(atspeed_session* s, int32_t n_users, const int32_t* prompts_dev,
                                         const int32_t* prompt_lens_host, int32_t gamma, int32_t* tokens_dev,
                                         float* scores_dev, atspeed_stats* stats, void* stream) {
    return bssd_batch_impl(s, n_users, prompts_dev, prompt_lens_host, gamma, nullptr, nullptr, nullptr, tokens_dev, scores_dev,
                           stats, true, stream);
}
