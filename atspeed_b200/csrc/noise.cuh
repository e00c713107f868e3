// Counter-based randomness for AtSpeed-R (sampling) mode: Philox4x32-10 keyed by the session seed.
//
// The reference draws from torch's generators at five sites per round (code/beamSD.py:72-74 draft multinomial,
// :304 bonus multinomial, :336 rand_like, :343 randperm on the CPU generator, :363 residual multinomial); a
// stateful generator would force a device->host round trip or a [rows*V] noise tensor per draw.  Here every draw is a
// pure function of (seed, stream, index): `stream` names the user / round / level / site, `index` the position in the
// reference's flat [n_prev * V] index space (or the draft-pick position for the acceptance test).  Tests fetch the same
// numbers through atspeed_noise_fill and replay them into the CPU oracle, so both sides consume identical noise.
#pragma once
#include <stdint.h>

namespace atspeed {

enum NoiseSite : uint32_t {
    SITE_DRAFT = 0,      // draft-step multinomial          (beamSD.py:72-74), index = parent_pos * V + token
    SITE_ACCEPT = 1,     // acceptance uniforms r_i          (beamSD.py:336),   index = draft pick position
    SITE_PERM = 2,       // "random K of the accepted"       (beamSD.py:343),   index = draft pick position
    SITE_RESIDUAL = 3,   // residual multinomial             (beamSD.py:363),   index = prev_pos * V + token
    SITE_BONUS = 4,      // bonus-level multinomial          (beamSD.py:304),   index = carried_row * V + token
    SITE_STEP = 5,       // plain target step / target_generate sampling (beamSD.py:72-74 on the target)
};

// stream id layout: [63:16] user sequence number, [15:8] round, [7:4] level, [3:0] site
__host__ __device__ __forceinline__ uint64_t noise_stream(uint64_t user_seq, uint32_t round, uint32_t level, uint32_t site) {
    return (user_seq << 16) | (static_cast<uint64_t>(round & 0xffu) << 8) | (static_cast<uint64_t>(level & 0xfu) << 4) |
           static_cast<uint64_t>(site & 0xfu);
}

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
    return static_cast<uint32_t>((static_cast<uint64_t>(a) * static_cast<uint64_t>(b)) >> 32);
}

// Philox4x32-10 (Salmon et al., SC'11): counter = (index, 0, stream_lo, stream_hi), key = (seed_lo, seed_hi); word 0.
__host__ __device__ __forceinline__ uint32_t philox_u32(uint64_t seed, uint64_t stream, uint32_t index) {
    uint32_t c0 = index, c1 = 0u, c2 = static_cast<uint32_t>(stream), c3 = static_cast<uint32_t>(stream >> 32);
    uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c0;
}

// uniform in (0, 1): 23 random bits + 0.5 ulp, exactly representable in fp32 (never 0, never 1)
__host__ __device__ __forceinline__ float u32_to_uniform(uint32_t x) {
    return (static_cast<float>(x >> 9) + 0.5f) * (1.0f / 8388608.0f);
}

#ifdef __CUDACC__
__device__ __forceinline__ float noise_uniform(uint64_t seed, uint64_t stream, uint32_t index) {
    return u32_to_uniform(philox_u32(seed, stream, index));
}
// Exp(1) noise for the exponential-race form of sampling without replacement:
// multinomial(p, n) == top-n of p / e with e ~ Exp(1) (what ATen's no-replacement path computes; SURVEY 7 hard part 4)
__device__ __forceinline__ float noise_exponential(uint64_t seed, uint64_t stream, uint32_t index) {
    return -logf(noise_uniform(seed, stream, index));
}
#endif

}  // namespace atspeed
