// Native restatement of numpy's legacy RandomState streams (MT19937), used on the
// host so that minibatch permutations, replay sample indices and exploration
// noise are bit-identical to the reference's numpy calls:
//   tonic/replays/segments.py:20,62   RandomState(seed).shuffle(arange(T*N))
//   tonic/replays/buffers.py:22,86    RandomState(seed).randint(total, size=B)
//   tonic/explorations/noisy.py:13,21,41  uniform(-1, 1, shape) / normal(size=shape)
// Algorithm: numpy/random/src/mt19937 + legacy-distributions (numpy 2.3.5, a
// third-party dependency of the reference that is not under /root/reference);
// pinned against numpy itself by tests/test_host_rng.py and the golden KATs.
#include <cmath>
#include <cstdint>
#include <cstdlib>

#include "../../include/tonic_b200.h"

struct TbRandomState {
    uint32_t mt[624];
    int pos;
    bool has_gauss;
    double gauss;
};

static void mt_seed(TbRandomState* s, uint32_t seed) {     // init_genrand
    s->mt[0] = seed;
    for (int i = 1; i < 624; ++i)
        s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
    s->pos = 624;
    s->has_gauss = false;
    s->gauss = 0.0;
}

static void mt_generate(TbRandomState* s) {
    uint32_t* mt = s->mt;
    const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, A = 0x9908b0dfu;
    int i;
    for (i = 0; i < 624 - 397; ++i) {
        const uint32_t y = (mt[i] & UPPER) | (mt[i + 1] & LOWER);
        mt[i] = mt[i + 397] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
    }
    for (; i < 623; ++i) {
        const uint32_t y = (mt[i] & UPPER) | (mt[i + 1] & LOWER);
        mt[i] = mt[i + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
    }
    const uint32_t y = (mt[623] & UPPER) | (mt[0] & LOWER);
    mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
    s->pos = 0;
}

static inline uint32_t next32(TbRandomState* s) {
    if (s->pos == 624) mt_generate(s);
    uint32_t y = s->mt[s->pos++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}
static inline uint64_t next64(TbRandomState* s) {
    const uint64_t hi = next32(s);
    return (hi << 32) | next32(s);
}
static inline double next_double(TbRandomState* s) {
    const int32_t a = (int32_t)(next32(s) >> 5), b = (int32_t)(next32(s) >> 6);
    return (a * 67108864.0 + b) / 9007199254740992.0;
}

// legacy random_interval / masked rejection sampling on [0, max]
static inline uint64_t bounded(TbRandomState* s, uint64_t max) {
    if (max == 0) return 0;
    uint64_t mask = max;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4;
    mask |= mask >> 8; mask |= mask >> 16; mask |= mask >> 32;
    uint64_t v;
    if (max <= 0xffffffffull) {
        while ((v = (next32(s) & mask)) > max) {}
    } else {
        while ((v = (next64(s) & mask)) > max) {}
    }
    return v;
}

extern "C" TbRandomState* tb_rs_create(uint32_t seed) {
    TbRandomState* s = (TbRandomState*)malloc(sizeof(TbRandomState));
    if (s) mt_seed(s, seed);
    return s;
}
extern "C" void tb_rs_destroy(TbRandomState* rs) { free(rs); }

extern "C" void tb_rs_shuffle_i64(TbRandomState* rs, int64_t* x, int64_t n) {
    for (int64_t i = n - 1; i >= 1; --i) {                  // _shuffle_raw
        const int64_t j = (int64_t)bounded(rs, (uint64_t)i);
        const int64_t t = x[i]; x[i] = x[j]; x[j] = t;
    }
}

extern "C" void tb_rs_randint(TbRandomState* rs, int64_t high, int64_t* out, int64_t n) {
    const uint64_t rng = (uint64_t)(high - 1);              // _rand_int64(0, high - 1), masked
    for (int64_t i = 0; i < n; ++i) {
        if (rng == 0xffffffffull) out[i] = (int64_t)next32(rs);
        else out[i] = (int64_t)bounded(rs, rng);
    }
}

extern "C" void tb_rs_uniform(TbRandomState* rs, double low, double high, double* out, int64_t n) {
    const double scale = high - low;
    for (int64_t i = 0; i < n; ++i) out[i] = low + scale * next_double(rs);
}

extern "C" void tb_rs_normal(TbRandomState* rs, double* out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) {                       // legacy_gauss (polar Box-Muller)
        if (rs->has_gauss) {
            rs->has_gauss = false;
            out[i] = rs->gauss;
            rs->gauss = 0.0;
            continue;
        }
        double f, x1, x2, r2;
        do {
            x1 = 2.0 * next_double(rs) - 1.0;
            x2 = 2.0 * next_double(rs) - 1.0;
            r2 = x1 * x1 + x2 * x2;
        } while (r2 >= 1.0 || r2 == 0.0);
        f = sqrt(-2.0 * log(r2) / r2);
        rs->gauss = f * x1;
        rs->has_gauss = true;
        out[i] = f * x2;
    }
}
