// psdr-b200: per-node arithmetic of the device LBVH build (pb_lbvh.cu), host-callable so that tests/native/lbvh_check.cu can run it on the CPU.
#pragma once
#include "pb_math.cuh"

namespace pb {

PB_HD int lbvh_clz(unsigned x) {
#ifdef __CUDA_ARCH__
    return __clz((int)x);
#else
    return x ? __builtin_clz(x) : 32;
#endif
}

PB_HD unsigned lbvh_expand_bits10(unsigned v) {   // 10 bits -> every third bit
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
// 30-bit Morton code of a point in the box [lo, lo + 1 / inv_ext]
PB_HD unsigned lbvh_morton30(float cx, float cy, float cz, float3 lo, float3 inv_ext) {
    const unsigned ux = (unsigned)fminf(fmaxf((cx - lo.x) * inv_ext.x * 1024.f, 0.f), 1023.f);
    const unsigned uy = (unsigned)fminf(fmaxf((cy - lo.y) * inv_ext.y * 1024.f, 0.f), 1023.f);
    const unsigned uz = (unsigned)fminf(fmaxf((cz - lo.z) * inv_ext.z * 1024.f, 0.f), 1023.f);
    return (lbvh_expand_bits10(ux) << 2) | (lbvh_expand_bits10(uy) << 1) | lbvh_expand_bits10(uz);
}

// length of the common prefix of the keys (code, index) at sorted positions i and j; -1 outside the array
PB_HD int lbvh_delta(const unsigned *codes, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const unsigned a = codes[i], b = codes[j];
    if (a == b) return 32 + lbvh_clz((unsigned)i ^ (unsigned)j);
    return lbvh_clz(a ^ b);
}

// Inner node i (of n - 1) of the radix tree over the sorted codes (Karras 2012): its range [first, last] of sorted slots and its two child
// references — >= 0: inner node; < 0: leaf over sorted slots, ~((first << 3) | (count - 1)), for ranges of at most leaf_max slots.
struct LbvhNode { int left, right, first, last; };
PB_HD LbvhNode lbvh_inner_node(const unsigned *codes, int n, int i, int leaf_max) {
    const int d = (lbvh_delta(codes, n, i, i + 1) - lbvh_delta(codes, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = lbvh_delta(codes, n, i, i - d);
    int lmax = 2;
    while (lbvh_delta(codes, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (lbvh_delta(codes, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = lbvh_delta(codes, n, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) / 2;
        if (lbvh_delta(codes, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + (d < 0 ? d : 0);
    LbvhNode r;
    r.first = i < j ? i : j; r.last = i < j ? j : i;
    const int cl = gamma - r.first + 1, cr = r.last - gamma;
    r.left = cl <= leaf_max ? ~((r.first << 3) | (cl - 1)) : gamma;
    r.right = cr <= leaf_max ? ~(((gamma + 1) << 3) | (cr - 1)) : gamma + 1;
    return r;
}

}  // namespace pb
