// Host-side epoch shuffle, bit-exact with NumPy's legacy
// RandomState.shuffle(arange(n)) that the reference uses
// (spotlight/torch_utils.py:46-47: Fisher-Yates from the end, j drawn by the
// masked-rejection bounded sampler on the model's MT19937 stream).
//
// NumPy spends ~30 ns per element here, almost all of it in cache misses on the
// random side of each swap.  The draws j_i do not depend on the array contents,
// so this version generates them a batch ahead, prefetches the cache lines they
// will touch, and only then performs the swaps in order -- the same sequence of
// swaps, therefore the same permutation and the same final generator state.
#include <stdint.h>
#include <string.h>

#include "../../include/spotlight_b200.h"

namespace {

constexpr int N = 624, M = 397;

struct Mt {
    uint32_t key[N];
    int pos;
    void twist() {
        int k = 0;
        for (; k < N - M; ++k) {
            const uint32_t y = (key[k] & 0x80000000u) | (key[k + 1] & 0x7fffffffu);
            key[k] = key[k + M] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        for (; k < N - 1; ++k) {
            const uint32_t y = (key[k] & 0x80000000u) | (key[k + 1] & 0x7fffffffu);
            key[k] = key[k + (M - N)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        const uint32_t y = (key[N - 1] & 0x80000000u) | (key[0] & 0x7fffffffu);
        key[N - 1] = key[M - 1] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        pos = 0;
    }
    inline uint32_t next() {
        if (pos >= N) twist();
        uint32_t y = key[pos++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
};

inline uint64_t mask_for(uint64_t r) {
    r |= r >> 1; r |= r >> 2; r |= r >> 4; r |= r >> 8; r |= r >> 16; r |= r >> 32;
    return r;
}

template <typename T>
void shuffle_impl(Mt& mt, int64_t n, T* x) {
    constexpr int BATCH = 64;
    int64_t js[2][BATCH];
    // fill batch b with draws for i = hi, hi-1, ..., and prefetch their lines
    auto draw = [&](int b, int64_t hi, int count) {
        for (int k = 0; k < count; ++k) {
            const uint64_t i = static_cast<uint64_t>(hi - k);
            const uint32_t mask = static_cast<uint32_t>(mask_for(i));
            uint32_t v;
            do { v = mt.next() & mask; } while (v > i);
            js[b][k] = v;
            __builtin_prefetch(x + v, 1, 0);
        }
    };
    int64_t i = n - 1;
    if (i < 1) return;
    int cur = 0;
    int cnt = static_cast<int>(i < BATCH ? i : BATCH);
    draw(cur, i, cnt);
    while (cnt > 0) {
        const int64_t nxt_hi = i - cnt;
        const int nxt_cnt = static_cast<int>(nxt_hi < 1 ? 0 : (nxt_hi < BATCH ? nxt_hi : BATCH));
        if (nxt_cnt > 0) draw(cur ^ 1, nxt_hi, nxt_cnt);
        for (int k = 0; k < cnt; ++k) {
            const int64_t a = i - k, b = js[cur][k];
            const T t = x[a]; x[a] = x[b]; x[b] = t;
        }
        i = nxt_hi; cnt = nxt_cnt; cur ^= 1;
    }
}

}  // namespace

extern "C" {

int slb_host_shuffle_order(uint32_t* key /* [624] in/out */, int32_t* pos /* in/out */, int64_t n,
                           int32_t elem_bytes, void* order_out) {
    if (!key || !pos || !order_out || n < 0 || (elem_bytes != 4 && elem_bytes != 8)) return SLB_EINVAL;
    if (elem_bytes == 4 && n >= (1ll << 31)) return SLB_EINVAL;
    if (n - 1 > 0xfffffffell) return SLB_EINVAL;           // legacy 32-bit draw path only
    Mt mt;
    memcpy(mt.key, key, sizeof(mt.key));
    mt.pos = *pos;
    if (elem_bytes == 4) {
        int32_t* x = static_cast<int32_t*>(order_out);
        for (int64_t k = 0; k < n; ++k) x[k] = static_cast<int32_t>(k);
        shuffle_impl(mt, n, x);
    } else {
        int64_t* x = static_cast<int64_t*>(order_out);
        for (int64_t k = 0; k < n; ++k) x[k] = k;
        shuffle_impl(mt, n, x);
    }
    memcpy(key, mt.key, sizeof(mt.key));
    *pos = mt.pos;
    return SLB_OK;
}

}  // extern "C"
