// Device-side prompt builder: the step in front of the hot path.
//
// The reference builds a test prompt per user on the host: SeqRecDataset._process_test_data (code/data.py:232-263) joins the
// code strings of the user's last <= 20 history items into the SFT template (code/data.py:19-20, :253-255), TestCollator
// (code/collator.py:50-75) tokenises it.  An item is exactly 4 code tokens whose vocabulary ids are known per item, the
// template text is a fixed token run, so the token ids of a prompt are a pure function of the user's history item ids:
//     [BOS] prefix  item_0 sep item_1 sep ... item_{h-1}  suffix  "Response:"
// (sep alternates between the one- and the two-piece tokenisation of ", ", as atspeed_b200/prompts.py lays it out).  One CTA
// per user writes its prompt straight into the concatenated int32 buffer atspeed_bssd_batch_device consumes; the host only
// needs the prompt LENGTHS, a closed form of the history length (prompts.prompt_len).  Integer work, bit-exact by test
// (tests/test_gpu_prompts.py against prompts.RecDataset.prompt_ids).
#include "../../include/atspeed.h"
#include "common.cuh"

namespace atspeed {

__global__ void build_prompts_kernel(const int* __restrict__ item_tok, const int* __restrict__ hist_items,
                                     const long long* __restrict__ hist_begin, const int* __restrict__ hist_len,
                                     const long long* __restrict__ out_off, atspeed_prompt_template tp, int* __restrict__ out) {
    const int u = blockIdx.x;
    const int h = hist_len[u];
    const int* hist = hist_items + hist_begin[u];
    int* dst = out + out_off[u];
    const int body0 = 1 + tp.n_prefix;
    // separators before item j: ceil(j/2) one-token + floor(j/2) two-token ones
    auto item_pos = [&](int j) { return body0 + 4 * j + ((j + 1) >> 1) * tp.n_sep_even + (j >> 1) * tp.n_sep_odd; };
    const int tail0 = h > 0 ? item_pos(h - 1) + 4 : body0;
    if (threadIdx.x == 0) dst[0] = tp.bos;
    for (int i = threadIdx.x; i < tp.n_prefix; i += blockDim.x) dst[1 + i] = tp.ids[i];
    for (int j = threadIdx.x; j < h; j += blockDim.x) {
        const int p = item_pos(j);
        const int* t = item_tok + static_cast<long long>(hist[j]) * 4;
        dst[p] = t[0]; dst[p + 1] = t[1]; dst[p + 2] = t[2]; dst[p + 3] = t[3];
        if (j + 1 < h) {
            const bool even = (j & 1) == 0;
            const int n = even ? tp.n_sep_even : tp.n_sep_odd;
            const int* sp = tp.ids + tp.n_prefix + tp.n_suffix + tp.n_resp + (even ? 0 : tp.n_sep_even);
            for (int k = 0; k < n; ++k) dst[p + 4 + k] = sp[k];
        }
    }
    for (int i = threadIdx.x; i < tp.n_suffix + tp.n_resp; i += blockDim.x) dst[tail0 + i] = tp.ids[tp.n_prefix + i];
}

}  // namespace atspeed

extern "C" int atspeed_build_prompts(const int32_t* item_tok_dev, const int32_t* hist_items_dev, const int64_t* hist_begin_dev,
                                     const int32_t* hist_len_dev, const int64_t* out_off_dev, int32_t n_users,
                                     const atspeed_prompt_template* tmpl, int32_t* prompts_dev, void* stream) {
    using namespace atspeed;
    ATS_CHECK_ARG(item_tok_dev && hist_items_dev && hist_begin_dev && hist_len_dev && out_off_dev && tmpl && prompts_dev,
                  "null argument");
    ATS_CHECK_ARG(n_users >= 1, "n_users=%d", n_users);
    ATS_CHECK_ARG(tmpl->n_prefix >= 0 && tmpl->n_suffix >= 0 && tmpl->n_resp >= 0 && tmpl->n_sep_even >= 0 && tmpl->n_sep_odd >= 0 &&
                      tmpl->n_prefix + tmpl->n_suffix + tmpl->n_resp + tmpl->n_sep_even + tmpl->n_sep_odd <= ATSPEED_PROMPT_TEMPLATE_IDS,
                  "prompt template does not fit %d ids", ATSPEED_PROMPT_TEMPLATE_IDS);
    build_prompts_kernel<<<n_users, 64, 0, static_cast<cudaStream_t>(stream)>>>(
        item_tok_dev, hist_items_dev, reinterpret_cast<const long long*>(hist_begin_dev), hist_len_dev,
        reinterpret_cast<const long long*>(out_off_dev), *tmpl, prompts_dev);
    ATS_LAUNCH_CHECK();
    return ATS_OK;
}
