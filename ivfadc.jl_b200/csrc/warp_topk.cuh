// Warp-distributed sorted candidate list: 32*R (value, payload) entries kept ascending across
// the lanes of one warp, entry e living in lane e / R, register e % R ("blocked" layout, so
// that an insertion moves exactly one register across a lane boundary: one shuffle per field).
#pragma once

#include "common.cuh"

namespace ivf {

template <typename T, typename P, int R> struct WarpList {
    T v[R];
    P p[R];

    __device__ __forceinline__ void init(P empty_payload) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            v[r] = Limits<T>::inf();
            p[r] = empty_payload;
        }
    }

    // Value of entry `e` (warp-uniform e), broadcast to all lanes.
    __device__ __forceinline__ T value_at(int e) const {
        const int kr = e % R;
        T x = v[0];
#pragma unroll
        for (int r = 1; r < R; ++r)
            if (r == kr) x = v[r];
        return __shfl_sync(0xffffffffu, x, e / R);
    }
    __device__ __forceinline__ P payload_at(int e) const {
        const int kr = e % R;
        P x = p[0];
#pragma unroll
        for (int r = 1; r < R; ++r)
            if (r == kr) x = p[r];
        return __shfl_sync(0xffffffffu, x, e / R);
    }

    // Insert (nv, np), warp-uniform arguments, called by all 32 lanes.  The new entry goes in
    // front of the first entry that is "greater"; with LEX = false that is `entry.v > nv` (the new
    // entry lands AFTER equal values: stable, what the reference's SortedMultiDict / stable
    // sortperm do when candidates arrive in scan order); with LEX = true ties on the value are
    // broken by the payload (entry.p > np), for merging lists whose arrival order is arbitrary.
    // The last entry of the list falls off.
    template <bool LEX> __device__ __forceinline__ void insert(T nv, P np) {
        const int lane = threadIdx.x & 31;
        T pv = __shfl_up_sync(0xffffffffu, v[R - 1], 1);
        P pp = __shfl_up_sync(0xffffffffu, p[R - 1], 1);
        bool pg = lane > 0 && greater<LEX>(pv, pp, nv, np);
#pragma unroll
        for (int r = R - 1; r >= 0; --r) {
            const bool g = greater<LEX>(v[r], p[r], nv, np);
            bool gprev;
            T sv;
            P sp;
            if (r > 0) {
                gprev = greater<LEX>(v[r - 1], p[r - 1], nv, np);
                sv = v[r - 1];
                sp = p[r - 1];
            } else {
                gprev = pg;
                sv = pv;
                sp = pp;
            }
            if (g) {
                v[r] = gprev ? sv : nv;
                p[r] = gprev ? sp : np;
            }
        }
    }

    template <bool LEX>
    __device__ __forceinline__ static bool greater(T ev, P ep, T nv, P np) {
        if (LEX) return ev > nv || (ev == nv && ep > np);
        return ev > nv;
    }
};

}  // namespace ivf
