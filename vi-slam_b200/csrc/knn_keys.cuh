// knn_keys.cuh — packed (distance, index) keys shared by the kNN kernels and the match filter.
// key = distance << 23 | index for Hamming (distance <= 256 fits 9 bits): unsigned order of keys ==
// (distance ascending, index ascending), which is cv::BFMatcher's result order (ties -> lowest train index).
#pragma once
#include <stdint.h>

constexpr uint32_t KEY_INF = 0xFFFFFFFFu;
constexpr int KEY_SHIFT = 23;
constexpr uint32_t KEY_IDX_MASK = (1u << KEY_SHIFT) - 1u;

// float-distance variant (L2): key = float bits << 32 | index (non-negative floats order like their bits)
constexpr unsigned long long KEY64_INF = 0xFFFFFFFFFFFFFFFFull;

template <typename K>
__device__ __forceinline__ void top2_insert(K& b0, K& b1, K k) {
    K hi = max(b0, k);
    b0 = min(b0, k);
    b1 = min(b1, hi);
}

// merge two sorted pairs (a0 <= a1), (o0 <= o1) holding disjoint candidates
template <typename K>
__device__ __forceinline__ void top2_merge(K& a0, K& a1, K o0, K o1) {
    K hi = max(a0, o0);
    K lo2 = min(a1, o1);
    a0 = min(a0, o0);
    a1 = min(hi, lo2);
}

// Concurrent top-2 on a global pair p[0] <= p[1] initialised to all-ones, keys unique:
// the loser of the first atomicMin is offered to the second slot.  The global minimum never loses and
// the second smallest key loses exactly once, so p ends as the two smallest keys regardless of order.
__device__ __forceinline__ void top2_publish_one(uint32_t* p, uint32_t k) {
    if (k == KEY_INF) return;
    uint32_t old = atomicMin(p, k);
    uint32_t loser = max(old, k);
    if (loser != KEY_INF) atomicMin(p + 1, loser);
}
__device__ __forceinline__ void top2_publish(uint32_t* p, uint32_t m0, uint32_t m1) {
    top2_publish_one(p, m0);
    top2_publish_one(p, m1);
}
__device__ __forceinline__ void top2_publish_one(unsigned long long* p, unsigned long long k) {
    if (k == KEY64_INF) return;
    unsigned long long old = atomicMin(p, k);
    unsigned long long loser = max(old, k);
    if (loser != KEY64_INF) atomicMin(p + 1, loser);
}
__device__ __forceinline__ void top2_publish(unsigned long long* p, unsigned long long m0, unsigned long long m1) {
    top2_publish_one(p, m0);
    top2_publish_one(p, m1);
}
