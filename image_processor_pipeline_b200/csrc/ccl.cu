// ccl.cu — keep the largest 8-connected component (see ccl.cuh for the design).
#include "ccl.cuh"

namespace ipp {

constexpr uint64_t kEven = 0x5555555555555555ull;

__device__ __forceinline__ uint32_t compress_even(uint64_t x) {
    x &= kEven;
    x = (x | (x >> 1)) & 0x3333333333333333ull;
    x = (x | (x >> 2)) & 0x0f0f0f0f0f0f0f0full;
    x = (x | (x >> 4)) & 0x00ff00ff00ff00ffull;
    x = (x | (x >> 8)) & 0x0000ffff0000ffffull;
    x = (x | (x >> 16)) & 0x00000000ffffffffull;
    return (uint32_t)x;
}

// pixels 64k .. 64k+63 of one mask row
__device__ __forceinline__ uint64_t load_px64(const uint32_t* row, int k, int wpr) {
    if ((wpr & 1) == 0) {                               // even row length (mask base is 256-byte aligned): rows are 8-byte aligned
        if (2 * k >= wpr) return 0ull;
        const uint2 v = *reinterpret_cast<const uint2*>(row + 2 * k);
        return (uint64_t)v.x | ((uint64_t)v.y << 32);
    }
    const uint32_t lo = (2 * k < wpr) ? row[2 * k] : 0u;
    const uint32_t hi = (2 * k + 1 < wpr) ? row[2 * k + 1] : 0u;
    return (uint64_t)lo | ((uint64_t)hi << 32);
}

// first block of the run that contains block gbx (row-local block column)
__device__ __forceinline__ int run_start(const uint32_t* crow, int gbx) {
    int k = gbx >> 5;
    const int i = gbx & 31;
    uint32_t m = ~crow[k] & (0xffffffffu >> (31 - i));
    while (m == 0u) {
        --k;
        m = ~crow[k];
    }
    return (k << 5) + 31 - __clz(m);
}

__device__ __forceinline__ int find_root(const int2* L, int x) {
    int p;
    while ((p = __ldcg(&L[x].x)) != x) x = p;
    return x;
}

// find with path halving (used while unions are in flight): every visited node is re-pointed to
// its grandparent with atomicMin — parents only ever decrease, so concurrent unions stay valid —
// which keeps the chains short for the later passes (area / keep), whose time is the length of
// the longest pointer chase.
__device__ __forceinline__ int find_root_halving(int2* L, int x) {
    int p = __ldcg(&L[x].x);
    while (p != x) {
        const int gp = __ldcg(&L[p].x);
        if (gp != p) atomicMin(&L[x].x, gp);
        x = p;
        p = gp;
    }
    return x;
}

__device__ __forceinline__ void unite(int2* L, int a, int b) {
    bool done;
    do {
        a = find_root_halving(L, a);
        b = find_root_halving(L, b);
        if (a < b) {
            const int old = atomicMin(&L[b].x, a);
            done = (old == b);
            b = old;
        } else if (b < a) {
            const int old = atomicMin(&L[a].x, b);
            done = (old == a);
            a = old;
        } else {
            done = true;
        }
    } while (!done);
}

struct GroupIdx {
    int img, by, k;
    bool ok;
};
// grid = (ceil(groups per image / (256 * kRep)), images): 32-bit index math, a CTA never straddles two images.
// Every thread visits kRep groups (rep = 0 .. kRep-1, 256 apart so that each visit of a warp is coalesced) and issues the
// first loads of all of them before it looks at any: most groups of a keyed frame are empty, and one thread per group
// spent its life waiting for a single load (the passes were bound by resident threads / latency, not by bandwidth).
// measured per pass (512 keyed 1080p frames): the scans over mostly empty groups gain from 2-4 visits per thread, the
// passes dominated by the union-find work of the big component want the threads (2 visits)
#ifndef IPP_CCL_R_INIT
#define IPP_CCL_R_INIT 4
#define IPP_CCL_R_MERGE 2
#define IPP_CCL_R_AREA 2
#define IPP_CCL_R_BEST 4
#define IPP_CCL_R_KEEP 2
#endif
constexpr int kRepInit = IPP_CCL_R_INIT, kRepMerge = IPP_CCL_R_MERGE, kRepArea = IPP_CCL_R_AREA, kRepBest = IPP_CCL_R_BEST, kRepKeep = IPP_CCL_R_KEEP;
template <int kRep>
__device__ __forceinline__ GroupIdx group_index(const CclGeom& g, int n, int rep) {
    GroupIdx r;
    const uint32_t rem = (blockIdx.x * kRep + rep) * blockDim.x + threadIdx.x;
    const uint32_t p = rem >> 5, l = rem & 31u;            // patch, lane inside the patch
    uint32_t q = p;                                        // p / pw
    if (g.pw > 1) {
        q = __umulhi(p, g.pw_inv);                         // pw_inv = floor(2^32 / pw) + 1: exact or one too large
        if (q * (uint32_t)g.pw > p) --q;
    }
    r.img = (int)blockIdx.y;
    r.by = (int)(q * kCclPatchH + l / kCclPatchW);
    r.k = (int)((p - q * (uint32_t)g.pw) * kCclPatchW + l % kCclPatchW);
    r.ok = r.by < g.bh && r.k < g.bww && r.img < n;
    return r;
}

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w > v ? w : v;
    }
    return v;
}

// pixels of `row` dilated by one to the left and right (8-neighbourhood seen from the row above/below),
// for the 64 pixels of group k; 0 when the row does not exist
__device__ __forceinline__ uint64_t dilated_px64(const uint32_t* row, int k, int wpr) {
    const uint64_t v = load_px64(row, k, wpr);
    const uint64_t prev = (k > 0) ? (uint64_t)(row[2 * k - 1] >> 31) : 0ull;
    const uint64_t next = (2 * k + 2 < wpr) ? (uint64_t)(row[2 * k + 2] & 1u) : 0ull;
    return v | (v << 1) | (v >> 1) | prev | (next << 63);
}

// ---- phase 0: per-image reduction state
__global__ void ccl_reset_kernel(const CclBuffers b, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    b.best[i] = 0ull;
    int32_t* acc = b.acc + 8 * i;
    acc[0] = INT_MAX; acc[1] = INT_MAX; acc[2] = 0; acc[3] = 0; acc[4] = 0; acc[5] = 0; acc[6] = 0; acc[7] = 0;
}

// ---- phase 1: block bits, horizontal connectivity, initialise run-start entries.
// ISOLATED RUNS — runs that lie inside one 32-block group and touch no foreground pixel of the pixel
// rows above and below — are complete components (a keyed frame has thousands of them: specks).
// They are finished right here: area = popcount, straight into the per-image argmax; their blocks are
// cleared from F so that no later phase spends a label entry, a pointer chase or an atomic on them.
// (F then means "block of a run that needs union-find"; a set pixel whose F bit is clear belongs to an
// isolated run, whose identity is its first block.)
__global__ void __launch_bounds__(256) ccl_init_kernel(const CclGeom g, const CclBuffers b, int n) {
    unsigned long long cand = 0ull;
    int nfg_iso = 0;
    uint64_t P0[kRepInit], P1[kRepInit];
#pragma unroll
    for (int rep = 0; rep < kRepInit; ++rep) {
        const GroupIdx t = group_index<kRepInit>(g, n, rep);
        P0[rep] = 0ull; P1[rep] = 0ull;
        if (t.ok) {
            const uint32_t* r0 = b.bits + (int64_t)t.img * g.bits_stride + (int64_t)(2 * t.by) * g.wpr;
            P0[rep] = load_px64(r0, t.k, g.wpr);
            if ((2 * t.by + 1) < g.h) P1[rep] = load_px64(r0 + g.wpr, t.k, g.wpr);
        }
    }
#pragma unroll
    for (int rep = 0; rep < kRepInit; ++rep) {
        const GroupIdx t = group_index<kRepInit>(g, n, rep);
        if (!t.ok) continue;
        const uint32_t* bits = b.bits + (int64_t)t.img * g.bits_stride;
        const uint32_t* r0 = bits + (int64_t)(2 * t.by) * g.wpr;
        const bool has1 = (2 * t.by + 1) < g.h;
        const uint32_t* r1 = r0 + g.wpr;
        const uint64_t p0 = P0[rep], p1 = P1[rep];
        const uint64_t any = p0 | p1;
        const uint32_t Lc = compress_even(any), Rc = compress_even(any >> 1);
        uint32_t F = Lc | Rc;
        uint32_t carry = 0;
        if (t.k > 0 && (Lc & 1u)) carry = (r0[2 * t.k - 1] | (has1 ? r1[2 * t.k - 1] : 0u)) >> 31;   // only bit 0 of C needs it
        const uint32_t C = ((Rc << 1) | carry) & Lc;
        const int64_t cf = (int64_t)t.img * g.cf_stride + (int64_t)t.by * g.bww + t.k;
        if (F) {
            // blocks with a pixel that touches the row above / below
            uint64_t touch = 0ull;
            if (t.by > 0) touch |= p0 & dilated_px64(r0 - g.wpr, t.k, g.wpr);
            if (2 * t.by + 2 < g.h) touch |= p1 & dilated_px64(r1 + g.wpr, t.k, g.wpr);
            const uint32_t X = compress_even(touch | (touch >> 1));
            // does the run that reaches block 31 continue in the next group ?
            uint32_t next_c0 = 0;
            if ((Rc >> 31) && 2 * t.k + 2 < g.wpr) next_c0 = (r0[2 * t.k + 2] | (has1 ? r1[2 * t.k + 2] : 0u)) & 1u;
            int2* L = b.la + (int64_t)t.img * g.lab_stride;
            const int base = t.by * g.bw + (t.k << 5);
            uint32_t S = F & ~C, iso = 0;
            while (S) {
                const int i = __ffs(S) - 1;
                S &= S - 1;
                const uint32_t above = (i == 31) ? 0u : (~C & (0xffffffffu << (i + 1)));
                const int e = above ? (__ffs(above) - 1) : 32;                 // end (exclusive) of the run inside the group
                const uint32_t rm = ((e == 32) ? 0xffffffffu : ((1u << e) - 1u)) & ~((1u << i) - 1u);
                if ((X & rm) == 0u && !(e == 32 && next_c0)) {
                    const uint64_t pm = ((e == 32) ? ~0ull : ((1ull << (2 * e)) - 1ull)) & ~((1ull << (2 * i)) - 1ull);
                    const int area = __popcll(p0 & pm) + __popcll(p1 & pm);
                    const unsigned long long key =
                        ((unsigned long long)(uint32_t)area << 32) | (unsigned long long)(0xffffffffu - (uint32_t)(base + i));
                    cand = key > cand ? key : cand;
                    nfg_iso += area;
                    iso |= rm;
                } else {
                    L[base + i] = make_int2(base + i, 0);
                }
            }
            F &= ~iso;
        }
        b.cbits[cf] = C;
        b.fbits[cf] = F;
    }
    // isolated runs: one argmax / pixel-count atomic per warp (blockIdx.y is the image: a warp never straddles two)
    const int img = (int)blockIdx.y;
    const unsigned long long m = warp_max_u64(cand);
    const int tot = __reduce_add_sync(0xffffffffu, nfg_iso);
    if ((threadIdx.x & 31) == 0 && tot && img < n) {
        if (m > __ldcg(&b.best[img])) atomicMax(&b.best[img], m);   // best only grows: a stale read can only cause a redundant atomic
        atomicAdd(&b.acc[8 * img + 4], tot);
    }
}

// ---- phase 2: unions between runs of adjacent block rows (once per contact segment)
__device__ __forceinline__ void ccl_merge_group(const CclGeom& g, const CclBuffers& b, const GroupIdx& t);
__global__ void __launch_bounds__(256) ccl_merge_kernel(const CclGeom g, const CclBuffers b, int n) {
    // F == 0: nothing here, or only isolated runs (finished by the init pass; by definition they touch nothing above)
    uint32_t Fv[kRepMerge];
#pragma unroll
    for (int rep = 0; rep < kRepMerge; ++rep) {
        const GroupIdx t = group_index<kRepMerge>(g, n, rep);
        Fv[rep] = (t.ok && t.by > 0) ? b.fbits[(int64_t)t.img * g.cf_stride + (int64_t)t.by * g.bww + t.k] : 0u;
    }
#pragma unroll
    for (int rep = 0; rep < kRepMerge; ++rep)
        if (Fv[rep]) ccl_merge_group(g, b, group_index<kRepMerge>(g, n, rep));
}
__device__ __forceinline__ void ccl_merge_group(const CclGeom& g, const CclBuffers& b, const GroupIdx& t) {
    const uint32_t* bits = b.bits + (int64_t)t.img * g.bits_stride;
    const uint32_t* rowT = bits + (int64_t)(2 * t.by) * g.wpr;
    const uint32_t* rowB = rowT - g.wpr;
    const uint64_t T = load_px64(rowT, t.k, g.wpr);
    if (T == 0ull) return;
    const uint64_t B = load_px64(rowB, t.k, g.wpr);
    const uint64_t bprev = (t.k > 0) ? (uint64_t)(rowB[2 * t.k - 1] >> 31) : 0ull;
    const uint64_t bnext = (2 * t.k + 2 < g.wpr) ? (uint64_t)(rowB[2 * t.k + 2] & 1u) : 0ull;
    if ((B | bprev | bnext) == 0ull) return;
    const uint64_t Tl = T & kEven, Tr = (T >> 1) & kEven, Bl = B & kEven, Br = (B >> 1) & kEven;
    const uint32_t cU = compress_even((Tl | Tr) & (Bl | Br));
    const uint32_t cUL = compress_even(Tl & ((Br << 2) | bprev));
    const uint32_t cUR = compress_even(Tr & ((Bl >> 2) | (bnext << 62)));
    if ((cU | cUL | cUR) == 0u) return;
    const uint32_t* crow = b.cbits + (int64_t)t.img * g.cf_stride + (int64_t)t.by * g.bww;
    const uint32_t* crowU = crow - g.bww;
    const uint32_t C = crow[t.k], Cu = crowU[t.k];
    const uint32_t cu_next0 = (t.k + 1 < g.bww) ? (crowU[t.k + 1] & 1u) : 0u;
    const uint32_t Cush = (Cu >> 1) | (cu_next0 << 31);
    uint32_t nU = cU & ~(C & ((cUR << 1) | (Cu & (cU << 1))));
    uint32_t nUL = cUL & ~(cU & Cu) & ~(C & (cU << 1));
    uint32_t nUR = cUR & ~(cU & Cush);
    int2* L = b.la + (int64_t)t.img * g.lab_stride;
    const int rowbase = t.by * g.bw, rowbaseU = rowbase - g.bw, gb0 = t.k << 5;
    while (nU) {
        const int i = __ffs(nU) - 1;
        nU &= nU - 1;
        unite(L, rowbase + run_start(crow, gb0 + i), rowbaseU + run_start(crowU, gb0 + i));
    }
    while (nUL) {
        const int i = __ffs(nUL) - 1;
        nUL &= nUL - 1;
        unite(L, rowbase + run_start(crow, gb0 + i), rowbaseU + run_start(crowU, gb0 + i - 1));
    }
    while (nUR) {
        const int i = __ffs(nUR) - 1;
        nUR &= nUR - 1;
        unite(L, rowbase + run_start(crow, gb0 + i), rowbaseU + run_start(crowU, gb0 + i + 1));
    }
}

// first zero of C strictly above bit i (32 if none): end (exclusive) of the segment that starts at i
__device__ __forceinline__ int segment_end(uint32_t C, int i) {
    const uint32_t above = (i == 31) ? 0u : (~C & (0xffffffffu << (i + 1)));
    return above ? (__ffs(above) - 1) : 32;
}

// ---- phase 3: per-segment pixel counts onto the root, compress run-start entries.
// Lanes of a warp cover neighbouring block groups, so most of their segments belong to the
// same component: counts are combined per root inside the warp (match.any + redux) and a
// single atomicAdd per distinct root and round reaches L2 — a big blob costs ~1 atomic per
// warp instead of one per run segment (same-address atomics serialise in L2).
__global__ void __launch_bounds__(256) ccl_area_kernel(const CclGeom g, const CclBuffers b, int n) {
    uint32_t Fv[kRepArea];
#pragma unroll
    for (int rep = 0; rep < kRepArea; ++rep) {
        const GroupIdx t = group_index<kRepArea>(g, n, rep);
        Fv[rep] = t.ok ? b.fbits[(int64_t)t.img * g.cf_stride + (int64_t)t.by * g.bww + t.k] : 0u;
    }
    int nfg = 0;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int rep = 0; rep < kRepArea; ++rep) {
        const uint32_t F = Fv[rep];
        if (!__any_sync(0xffffffffu, F != 0u)) continue;
        const GroupIdx t = group_index<kRepArea>(g, n, rep);
        uint32_t starts = 0, C = 0;
        uint64_t p0 = 0, p1 = 0;
        const uint32_t* crow = nullptr;
        int2* L = nullptr;
        int rowbase = 0, gb0 = 0;
        if (F) {
            const int64_t cf = (int64_t)t.img * g.cf_stride + (int64_t)t.by * g.bww;
            C = b.cbits[cf + t.k];
            crow = b.cbits + cf;
            const uint32_t* bits = b.bits + (int64_t)t.img * g.bits_stride;
            const uint32_t* r0 = bits + (int64_t)(2 * t.by) * g.wpr;
            p0 = load_px64(r0, t.k, g.wpr);
            p1 = ((2 * t.by + 1) < g.h) ? load_px64(r0 + g.wpr, t.k, g.wpr) : 0ull;
            L = b.la + (int64_t)t.img * g.lab_stride;
            rowbase = t.by * g.bw;
            gb0 = t.k << 5;
            starts = F & (~C | 1u);
        }
        while (__any_sync(0xffffffffu, starts != 0u)) {
            long long key = -1 - (long long)lane;          // unique per lane when it has nothing to add
            int cnt = 0;
            int32_t* slot = nullptr;
            if (starts) {
                const int i = __ffs(starts) - 1;
                starts &= starts - 1;
                const int e = segment_end(C, i);
                const uint64_t pm = ((e == 32) ? ~0ull : ((1ull << (2 * e)) - 1ull)) & ~((1ull << (2 * i)) - 1ull);
                cnt = __popcll(p0 & pm) + __popcll(p1 & pm);
                const bool own = ((C >> i) & 1u) == 0u;
                const int s = own ? (gb0 + i) : run_start(crow, gb0 + i);
                const int root = find_root(L, rowbase + s);
                if (own) L[rowbase + gb0 + i].x = root;
                slot = &L[root].y;
                key = (long long)(uintptr_t)slot;
                nfg += cnt;
            }
            // neighbouring groups mostly sit in ONE component (the object): when every active lane has the same root a
            // single reduction and atomic do; only mixed warps pay for match.any
            const uint32_t active = __ballot_sync(0xffffffffu, slot != nullptr);
            const long long key0 = __shfl_sync(0xffffffffu, key, __ffs(active) - 1);
            if (__all_sync(0xffffffffu, slot == nullptr || key == key0)) {
                const int sum = __reduce_add_sync(0xffffffffu, cnt);
                if (slot && lane == __ffs(active) - 1) atomicAdd(slot, sum);
            } else {
                const uint32_t peers = __match_any_sync(0xffffffffu, key);
                const int sum = __reduce_add_sync(peers, cnt);
                if (slot && lane == __ffs(peers) - 1) atomicAdd(slot, sum);
            }
        }
    }
    // foreground pixel count per image: one atomic per warp (blockIdx.y is the image)
    const int tot = __reduce_add_sync(0xffffffffu, nfg);
    if (lane == 0 && tot) atomicAdd(&b.acc[8 * (int)blockIdx.y + 4], tot);
}

// ---- phase 4: argmax over roots of (area, smallest block index); one atomic per warp
__global__ void __launch_bounds__(256) ccl_best_kernel(const CclGeom g, const CclBuffers b, int n) {
    unsigned long long best = 0ull;
    uint32_t Sv[kRepBest];
#pragma unroll
    for (int rep = 0; rep < kRepBest; ++rep) {
        const GroupIdx t = group_index<kRepBest>(g, n, rep);
        Sv[rep] = 0u;
        if (t.ok) {
            const int64_t cf = (int64_t)t.img * g.cf_stride + (int64_t)t.by * g.bww + t.k;
            const uint32_t F = b.fbits[cf];
            if (F) Sv[rep] = F & ~b.cbits[cf];
        }
    }
#pragma unroll
    for (int rep = 0; rep < kRepBest; ++rep) {
        uint32_t S = Sv[rep];
        if (!S) continue;
        const GroupIdx t = group_index<kRepBest>(g, n, rep);
        const int2* L = b.la + (int64_t)t.img * g.lab_stride;
        const int base = t.by * g.bw + (t.k << 5);
        while (S) {
            const int i = __ffs(S) - 1;
            S &= S - 1;
            const int bi = base + i;
            const int2 la = L[bi];
            if (la.x == bi) {
                const unsigned long long key =
                    ((unsigned long long)(uint32_t)la.y << 32) | (unsigned long long)(0xffffffffu - (uint32_t)bi);
                best = key > best ? key : best;
            }
        }
    }
    const unsigned long long m = warp_max_u64(best);
    if ((threadIdx.x & 31) == 0 && m) atomicMax(&b.best[(int)blockIdx.y], m);
}

// each of the 32 low bits doubled: block bit i -> pixel bits 2i, 2i+1
__device__ __forceinline__ uint64_t spread_pairs(uint32_t v) {
    uint64_t x = v;
    x = (x | (x << 16)) & 0x0000ffff0000ffffull;
    x = (x | (x << 8)) & 0x00ff00ff00ff00ffull;
    x = (x | (x << 4)) & 0x0f0f0f0f0f0f0f0full;
    x = (x | (x << 2)) & 0x3333333333333333ull;
    x = (x | (x << 1)) & kEven;
    return x | (x << 1);
}

// ---- phase 5: keep-mask (in place) + bounding box of the kept pixels; one thread per 32-block group
// (64 pixels x 2 rows).  Run segments are classified once for both pixel rows; isolated runs (pixels
// whose F bit is clear) need no label lookup at all.
__global__ void __launch_bounds__(256) ccl_keep_kernel(const CclGeom g, const CclBuffers b, int n) {
    int minx = INT_MAX, miny = INT_MAX, maxx1 = 0, maxy1 = 0;
    uint64_t P0[kRepKeep], P1[kRepKeep];
#pragma unroll
    for (int rep = 0; rep < kRepKeep; ++rep) {
        const GroupIdx t = group_index<kRepKeep>(g, n, rep);
        P0[rep] = 0ull; P1[rep] = 0ull;
        if (t.ok) {
            const uint32_t* r0 = b.bits + (int64_t)t.img * g.bits_stride + (int64_t)(2 * t.by) * g.wpr;
            P0[rep] = load_px64(r0, t.k, g.wpr);
            if ((2 * t.by + 1) < g.h) P1[rep] = load_px64(r0 + g.wpr, t.k, g.wpr);
        }
    }
#pragma unroll
    for (int rep = 0; rep < kRepKeep; ++rep) {
        const uint64_t p0 = P0[rep], p1 = P1[rep];
        const uint64_t any = p0 | p1;
        if (any) {
            const GroupIdx t = group_index<kRepKeep>(g, n, rep);
            uint32_t* bits = b.bits + (int64_t)t.img * g.bits_stride;
            uint32_t* r0 = bits + (int64_t)(2 * t.by) * g.wpr;
            uint32_t* r1 = r0 + g.wpr;
            const unsigned long long bk = b.best[t.img];
            const int best_root = (int)(0xffffffffu - (uint32_t)(bk & 0xffffffffull));
            const int64_t cf = (int64_t)t.img * g.cf_stride + (int64_t)t.by * g.bww;
            const uint32_t* crow = b.cbits + cf;
            const uint32_t C = crow[t.k], F = b.fbits[cf + t.k];
            const uint32_t Fall = compress_even(any) | compress_even(any >> 1);
            const int2* L = b.la + (int64_t)t.img * g.lab_stride;
            const int rowbase = t.by * g.bw, gb0 = t.k << 5;
            uint32_t starts = Fall & (~C | 1u), keep = 0;
            // only isolated runs here and the winner is not one of them: nothing is kept, no need to walk the runs
            if (F == 0u && (unsigned)(best_root - (rowbase + gb0)) >= 32u) starts = 0u;
            while (starts) {
                const int i = __ffs(starts) - 1;
                starts &= starts - 1;
                const int e = segment_end(C, i);
                const uint32_t sm = ((e == 32) ? 0xffffffffu : ((1u << e) - 1u)) & ~((1u << i) - 1u);
                bool kept;
                if (((F >> i) & 1u) == 0u) {
                    kept = (rowbase + gb0 + i) == best_root;                  // isolated run: its first block is its identity
                } else {
                    const int s0 = (((C >> i) & 1u) == 0u) ? (gb0 + i) : run_start(crow, gb0 + i);
                    kept = L[rowbase + s0].x == best_root;
                }
                if (kept) keep |= sm;
            }
            const uint64_t km = spread_pairs(keep);
            const uint64_t k0 = p0 & km, k1 = p1 & km;
            if (k0 != p0) {
                r0[2 * t.k] = (uint32_t)k0;
                if (2 * t.k + 1 < g.wpr) r0[2 * t.k + 1] = (uint32_t)(k0 >> 32);
            }
            if (k1 != p1) {
                r1[2 * t.k] = (uint32_t)k1;
                if (2 * t.k + 1 < g.wpr) r1[2 * t.k + 1] = (uint32_t)(k1 >> 32);
            }
            const uint64_t ka = k0 | k1;
            if (ka) {
                const int x0 = 64 * t.k + __ffsll((long long)ka) - 1, x1 = 64 * t.k + 64 - __clzll((long long)ka);
                minx = min(minx, x0);
                maxx1 = max(maxx1, x1);
                miny = min(miny, k0 ? 2 * t.by : 2 * t.by + 1);
                maxy1 = max(maxy1, k1 ? 2 * t.by + 2 : 2 * t.by + 1);
            }
        }
    }
    // reduce per warp (blockIdx.y is the image), then atomics
    minx = warp_min(minx); miny = warp_min(miny); maxx1 = warp_max(maxx1); maxy1 = warp_max(maxy1);
    if ((threadIdx.x & 31) == 0 && maxx1 > 0) {
        int32_t* acc = b.acc + 8 * (int)blockIdx.y;
        atomicMin(&acc[0], minx); atomicMin(&acc[1], miny); atomicMax(&acc[2], maxx1); atomicMax(&acc[3], maxy1);
    }
}

int ccl_run(const CclGeom& g, const CclBuffers& b, int n, cudaStream_t st) {
    const int64_t groups = (int64_t)g.npatch * 32;          // threads: one per group slot of every warp patch
    auto grid = [&](int rep) { return dim3((unsigned)ceil_div(groups, 256 * rep), (unsigned)n); };
    { ipp::LaunchTimer _lt("ccl_reset_kernel", st); ccl_reset_kernel<<<ceil_div(n, 256), 256, 0, st>>>(b, n); }
    IPP_LAUNCH_CHECK();
    { ipp::LaunchTimer _lt("ccl_init_kernel", st); ccl_init_kernel<<<grid(kRepInit), 256, 0, st>>>(g, b, n); }
    IPP_LAUNCH_CHECK();
    { ipp::LaunchTimer _lt("ccl_merge_kernel", st); ccl_merge_kernel<<<grid(kRepMerge), 256, 0, st>>>(g, b, n); }
    IPP_LAUNCH_CHECK();
    { ipp::LaunchTimer _lt("ccl_area_kernel", st); ccl_area_kernel<<<grid(kRepArea), 256, 0, st>>>(g, b, n); }
    IPP_LAUNCH_CHECK();
    { ipp::LaunchTimer _lt("ccl_best_kernel", st); ccl_best_kernel<<<grid(kRepBest), 256, 0, st>>>(g, b, n); }
    IPP_LAUNCH_CHECK();
    { ipp::LaunchTimer _lt("ccl_keep_kernel", st); ccl_keep_kernel<<<grid(kRepKeep), 256, 0, st>>>(g, b, n); }
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}

// ================================================================== stand-alone op
// alpha > 1 -> pixel mask.  One warp per 1024 pixels of a row: lane l tests pixels
// l, l+32, ... so every ballot is one finished mask word and the byte loads of a warp
// cover contiguous memory.
__global__ void __launch_bounds__(256) alpha_to_bits_kernel(const uint8_t* __restrict__ img, int64_t pitch, int64_t img_stride,
                                                            const CclGeom g, uint32_t* bits, int n) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int wpr32 = (g.wpr + 31) / 32;                    // warps per row
    const int64_t per = (int64_t)g.h * wpr32;
    if (warp >= per * n) return;
    const int im = (int)(warp / per);
    const int rem = (int)(warp - (int64_t)im * per);
    const int y = rem / wpr32, seg = rem - y * wpr32;
    const uint8_t* row = img + (int64_t)im * img_stride + (int64_t)y * pitch;
    uint32_t* out = bits + (int64_t)im * g.bits_stride + (int64_t)y * g.wpr;
    const int w0 = seg * 32;
    uint32_t mine = 0;
    for (int j = 0; j < 32 && w0 + j < g.wpr; ++j) {
        const int x = (w0 + j) * 32 + lane;
        const bool fg = (x < g.w) && (row[(int64_t)x * 4 + 3] > 1);
        const uint32_t word = __ballot_sync(0xffffffffu, fg);
        if (lane == j) mine = word;
    }
    if (w0 + lane < g.wpr) out[w0 + lane] = mine;
}

// alpha[not kept] = 0 when any foreground exists; bbox of alpha != 0 afterwards
__global__ void __launch_bounds__(256) apply_keep_kernel(uint8_t* img, int64_t pitch, int64_t img_stride, const CclGeom g,
                                                         const CclBuffers b, int n) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t per = (int64_t)g.h * g.wpr;
    if (warp >= per * n) return;
    const int im = (int)(warp / per);
    const int rem = (int)(warp - (int64_t)im * per);
    const int y = rem / g.wpr, kx = rem - y * g.wpr;
    const int x = kx * 32 + lane;
    const bool any_fg = b.acc[8 * im + 4] > 0;
    const uint32_t keep = b.bits[(int64_t)im * g.bits_stride + (int64_t)y * g.wpr + kx];
    bool nz = false;
    if (x < g.w) {
        uint8_t* pa = img + (int64_t)im * img_stride + (int64_t)y * pitch + (int64_t)x * 4 + 3;
        const uint8_t a = *pa;
        if (any_fg) {
            const bool kept = (keep >> lane) & 1u;
            if (!kept && a != 0) *pa = 0;
            nz = kept;            // kept pixels have alpha > 1
        } else {
            nz = a != 0;          // no component: image unchanged (alpha == 1 pixels count)
        }
    }
    const uint32_t m = __ballot_sync(0xffffffffu, nz);
    if (lane == 0 && m && !any_fg) {
        int32_t* acc = b.acc + 8 * im;
        atomicMin(&acc[0], 32 * kx + __ffs(m) - 1);
        atomicMax(&acc[2], 32 * kx + 32 - __clz(m));
        atomicMin(&acc[1], y);
        atomicMax(&acc[3], y + 1);
    }
}

__global__ void ccl_info_kernel(const CclBuffers b, int32_t* info, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t* acc = b.acc + 8 * i;
    int32_t* o = info + 8 * i;
    const bool any = acc[2] > 0;
    o[0] = any ? acc[0] : 0;
    o[1] = any ? acc[1] : 0;
    o[2] = any ? acc[2] - acc[0] : 0;
    o[3] = any ? acc[3] - acc[1] : 0;
    o[4] = (int32_t)(b.best[i] >> 32);
    o[5] = acc[4];
    o[6] = 0;
    o[7] = 0;
}

}  // namespace ipp

extern "C" {

int64_t ipp_keep_largest_workspace_bytes(int w, int h, int n) {
    if (w <= 0 || h <= 0 || n <= 0) return 0;
    return ipp::ccl_workspace_bytes(ipp::make_ccl_geom(w, h), n);
}

int ipp_keep_largest_cc_u8(uint8_t* bgra, int64_t pitch, int w, int h,
                           int32_t* info_out, void* workspace, int64_t workspace_bytes,
                           int n, int64_t image_stride, void* stream) {
    using namespace ipp;
    cudaStream_t st = (cudaStream_t)stream;
    IPP_REQUIRE(bgra && info_out, IPP_ERR_BAD_ARG, "ipp_keep_largest_cc_u8: NULL pointer");
    IPP_REQUIRE(w > 0 && h > 0 && n >= 1 && pitch >= (int64_t)w * 4, IPP_ERR_BAD_SHAPE,
                "ipp_keep_largest_cc_u8: bad geometry %dx%d n=%d", w, h, n);
    IPP_REQUIRE((int64_t)((w + 1) / 2) * ((h + 1) / 2) < (1LL << 31), IPP_ERR_UNSUPPORTED, "image too large");
    const CclGeom g = make_ccl_geom(w, h);
    IPP_REQUIRE(workspace && workspace_bytes >= ccl_workspace_bytes(g, n) && (((uintptr_t)workspace) & 255) == 0,
                IPP_ERR_WORKSPACE, "ipp_keep_largest_cc_u8: workspace too small or not 256-byte aligned (need %lld bytes)",
                (long long)ccl_workspace_bytes(g, n));
    const CclBuffers b = carve_ccl(workspace, g, n);
    {
        const int wpr32 = (g.wpr + 31) / 32;
        const int64_t warps = (int64_t)g.h * wpr32 * n;
        { ipp::LaunchTimer _lt("alpha_to_bits_kernel", st); alpha_to_bits_kernel<<<ceil_div(warps * 32, 256), 256, 0, st>>>(bgra, pitch, image_stride, g, b.bits, n); }
        IPP_LAUNCH_CHECK();
    }
    const int rc = ccl_run(g, b, n, st);
    if (rc != IPP_OK) return rc;
    {
        const int64_t warps = (int64_t)g.h * g.wpr * n;
        { ipp::LaunchTimer _lt("apply_keep_kernel", st); apply_keep_kernel<<<ceil_div(warps * 32, 256), 256, 0, st>>>(bgra, pitch, image_stride, g, b, n); }
        IPP_LAUNCH_CHECK();
    }
    { ipp::LaunchTimer _lt("ccl_info_kernel", st); ccl_info_kernel<<<ceil_div(n, 128), 128, 0, st>>>(b, info_out, n); }
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}

}  // extern "C"
