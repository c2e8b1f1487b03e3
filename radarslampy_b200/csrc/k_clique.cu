// k_clique.cu — distance-consistency graph + the reference's "first largest maximal clique".
//
// Replaces outlierRejection.rejectOutliers (outlierRejection.py:16-95):
//   cdist / |d_prev - d_new| <= thr adjacency (:49-58)          -> k_adjacency (fp64, no FMA)
//   nx.find_cliques + "first strictly larger clique" (:63-78)   -> k_clique
//
// Parity contract.  Maximum cliques tie on real data and the reference keeps the first one
// networkx yields, so the kernel reproduces networkx 3.6's pivoting Bron–Kerbosch *in the
// iteration order of CPython 3.12 sets* (restated in oracle/c/oracle_c.c).
// A set of small ints is a hash table whose slot of key k is k & mask (+ probing); when the
// table is at least as large as the node count every key sits in its own slot, i.e. the set
// iterates in ascending order and is fully described by a BITSET.  Only sets whose table is
// smaller than K need their slot layout emulated (an explicit int16 table of <= Kpad/2 slots).
// Each set is therefore { bitset | mask, fill, used, finger | small table }.
//
// Mapping: one warp per frame pair.  Bitset algebra, popcounts, ordered compaction and the
// pivot arg-max run across the 32 lanes; the collision-order emulation of small tables is
// inherently sequential and is done by lane 0.  The current search frame lives in shared
// memory; parents are pushed to a per-pair stack in global memory.
// Search-tree children that cannot beat the best clique so far are skipped (order-safe: the
// parent's state does not depend on whether a child was descended).
#include <stdlib.h>

#include "common.cuh"

#define EMPTY_SLOT (-1)
#define DUMMY_SLOT (-2)

struct CliqueGeom {
    int Kpad;    // padded capacity (multiple of 32, power of two)
    int NW;      // 32-bit words per bitset
    int TABN;    // int16 slots in an explicit table (Kpad / 2)
    int SW;      // 32-bit words per set = NW + 4 + TABN / 2
    int SEQCAP;  // max degree of a node whose adjacency set has a table smaller than Kpad
    int RS;      // row stride (words) of the adjacency rows staged in shared memory (odd: no bank conflicts)
};

__host__ __device__ inline int growth_size(int n) {
    // table size of a set grown by n successive adds from empty (CPython set_add_entry:
    // resize to used*4 when fill*5 >= mask*3)
    return n <= 4 ? 8 : n <= 18 ? 32 : n <= 76 ? 128 : n <= 306 ? 512 : n <= 1228 ? 2048 : 8192;
}

struct SetRef {
    uint32_t* w;  // base
    int NW;
    __device__ uint32_t* bits() const { return w; }
    __device__ int& mask() const { return ((int*)(w + NW))[0]; }
    __device__ int& fill() const { return ((int*)(w + NW))[1]; }
    __device__ int& used() const { return ((int*)(w + NW))[2]; }
    __device__ int& finger() const { return ((int*)(w + NW))[3]; }
    __device__ int16_t* tab() const { return (int16_t*)(w + NW + 4); }
};

#define FULL 0xffffffffu

// ---- sequential emulation helpers (lane 0 only; fallback for sequences longer than a warp) ----
__device__ void tab_insert_clean(int16_t* t, int mask, int key) {
    unsigned perturb = (unsigned)key;
    unsigned i = (unsigned)key & mask;
    for (;;) {
        if (t[i] == EMPTY_SLOT) { t[i] = (int16_t)key; return; }
        if (i + 9 <= (unsigned)mask) {
            for (int j = 1; j <= 9; ++j)
                if (t[i + j] == EMPTY_SLOT) { t[i + j] = (int16_t)key; return; }
        }
        perturb >>= 5;
        i = (i * 5 + 1 + perturb) & mask;
    }
}

__device__ __noinline__ void tab_build_by_adds_seq(int16_t* t, const int16_t* seq, int n, int16_t* tmp) {
    int mask = 7, fill = 0;
    for (int i = 0; i < 8; ++i) t[i] = EMPTY_SLOT;
    for (int s = 0; s < n; ++s) {
        tab_insert_clean(t, mask, seq[s]);
        ++fill;
        if (fill * 5 >= mask * 3) {
            int newsize = 8;
            while (newsize <= fill * 4) newsize <<= 1;
            int m = 0;
            for (int i = 0; i <= mask; ++i) if (t[i] >= 0) tmp[m++] = t[i];
            for (int i = 0; i < newsize; ++i) t[i] = EMPTY_SLOT;
            mask = newsize - 1;
            for (int i = 0; i < m; ++i) tab_insert_clean(t, mask, tmp[i]);
        }
    }
}

// ---- warp-parallel open-addressing insertion ----------------------------------------------
// Lane l holds the l-th key of an insertion sequence (l < n <= 32) into an EMPTY table.
// Sequential insertion places key j at the first slot of its probe sequence that no key
// i < j occupies.  That assignment is the unique fix-point of "every key claims its current
// probe slot; the lowest insertion index wins a contested slot; losers (and evicted keys)
// advance along their own probe sequence" — proved by induction on j: a slot that ever
// rejected j stays held by some key < j.  So all keys probe concurrently and only genuine
// collisions cost extra rounds.  Probe sequence = CPython set_insert_clean: i, i+1..i+9
// (only if i+9 <= mask), then i = (5i + 1 + (perturb >>= 5)) & mask.
__device__ __forceinline__ int par_insert(int key, bool active, int mask, int lane) {
    unsigned i = (unsigned)key & (unsigned)mask, j = 0, perturb = (unsigned)key;
    const int nbits = 32 - __clz(mask);            // mask = table size - 1 (warp-uniform)
    for (;;) {
        const int slot = (int)(i + j);
        // lanes probing the same slot, from one ballot per slot bit (slot <= mask): __match_any_sync gives the same
        // groups but was 20 % of the kernel's stall samples (ncu r01i, k_clique.cu:104)
        unsigned same = __ballot_sync(FULL, active);
        for (int b = 0; b < nbits; ++b) {
            const unsigned v = __ballot_sync(FULL, (slot >> b) & 1);
            same &= ((slot >> b) & 1) ? v : ~v;
        }
        const bool lose = active && (same & ((1u << lane) - 1u)) != 0u;
        if (!__any_sync(FULL, lose)) return slot;
        if (lose) {
            if (i + 9 <= (unsigned)mask && j < 9) ++j;
            else { perturb >>= 5; i = (i * 5 + 1 + perturb) & (unsigned)mask; j = 0; }
        }
    }
}

__device__ __forceinline__ void tab_clear(int16_t* t, int size, int lane) {
    uint32_t* t32 = reinterpret_cast<uint32_t*>(t);       // tables are 4-byte aligned, size >= 8
    for (int i = lane; i < (size >> 1); i += 32) t32[i] = 0xFFFFFFFFu;   // EMPTY_SLOT == -1
}

// Same fix-point for up to 96 keys (three per lane: key index j = lane + 32 r) into an empty
// table of at most 128 slots: contested slots are arbitrated through `own` (one word per
// slot, atomicMin of the insertion index) instead of a warp match.
#define PAR_KPL 3
__device__ __noinline__ void par_insert_multi(int16_t* t, int mask, const int16_t* seq, int n, uint32_t* own, int lane) {
    const int size = mask + 1;
    for (int i = lane; i < size; i += 32) own[i] = 0xFFFFFFFFu;
    unsigned pi[PAR_KPL], pj[PAR_KPL], pp[PAR_KPL];
    int key[PAR_KPL];
#pragma unroll
    for (int r = 0; r < PAR_KPL; ++r) {
        const int j = lane + 32 * r;
        key[r] = j < n ? seq[j] : 0;
        pi[r] = (unsigned)key[r] & (unsigned)mask; pj[r] = 0; pp[r] = (unsigned)key[r];
    }
    __syncwarp();
    for (;;) {
#pragma unroll
        for (int r = 0; r < PAR_KPL; ++r)
            if (lane + 32 * r < n) atomicMin(&own[pi[r] + pj[r]], (unsigned)(lane + 32 * r));
        __syncwarp();
        bool lose = false;
#pragma unroll
        for (int r = 0; r < PAR_KPL; ++r) {
            if (lane + 32 * r < n && own[pi[r] + pj[r]] != (unsigned)(lane + 32 * r)) {
                lose = true;
                if (pi[r] + 9 <= (unsigned)mask && pj[r] < 9) ++pj[r];
                else { pp[r] >>= 5; pi[r] = (pi[r] * 5 + 1 + pp[r]) & (unsigned)mask; pj[r] = 0; }
            }
        }
        if (!__any_sync(FULL, lose)) break;
    }
#pragma unroll
    for (int r = 0; r < PAR_KPL; ++r)
        if (lane + 32 * r < n) t[pi[r] + pj[r]] = (int16_t)key[r];
    __syncwarp();
}

// rank of this lane's slot among the first `m` lanes (slot order of a table of <= 32 slots)
__device__ __forceinline__ int slot_rank(int slot, int m, int lane) {
    const unsigned occ = __reduce_or_sync(FULL, lane < m ? (1u << slot) : 0u);
    return lane < m ? __popc(occ & ((1u << slot) - 1u)) : lane;
}

// Explicit table of a FRESH set grown by n successive adds of seq[0..n) (no dummies, no
// duplicates), CPython growth rule: 8 slots; after the 5th add -> 32 (old entries re-inserted
// in SLOT order); after the 19th -> 128; after the 77th -> 512; ...
__device__ __noinline__ void tab_build_by_adds(int16_t* t, const int16_t* seq, int n, int16_t* tmp, uint32_t* own, int lane) {
    if (n > 76) {
        if (lane == 0) tab_build_by_adds_seq(t, seq, n, tmp);
        __syncwarp();
        return;
    }
    const int size = n <= 4 ? 8 : (n <= 18 ? 32 : 128);
    tab_clear(t, size, lane);
    int key = lane < n ? seq[lane] : 0;
    int slot;
    if (n <= 4) {
        slot = par_insert(key, lane < n, 7, lane);
    } else {
        int s = par_insert(key, lane < 5, 7, lane);
        int rank = slot_rank(s, 5, lane);
        __syncwarp();
        if (lane < n) tmp[rank] = (int16_t)key;
        __syncwarp();
        key = lane < n ? tmp[lane] : 0;
        const int m = n < 19 ? n : 19;
        s = par_insert(key, lane < m, 31, lane);
        if (n < 19) slot = s;
        else {
            rank = slot_rank(s, 19, lane);
            __syncwarp();
            if (lane < n) tmp[rank] = (int16_t)key;
            __syncwarp();
            if (n > 32) {
                // insertion order into the 128-slot table: the 19 re-inserted keys, then seq[19..n)
                for (int j = 32 + lane; j < n; j += 32) tmp[j] = seq[j];
                __syncwarp();
                par_insert_multi(t, 127, tmp, n, own, lane);
                return;
            }
            key = lane < n ? tmp[lane] : 0;
            slot = par_insert(key, lane < n, 127, lane);
        }
    }
    __syncwarp();
    if (lane < n) t[slot] = (int16_t)key;
    __syncwarp();
}

// clean re-insertion of seq[0..m) (in that order) into an empty table with `mask`
__device__ __noinline__ void tab_reinsert(int16_t* t, int mask, const int16_t* seq, int m, uint32_t* own, int lane) {
    tab_clear(t, mask + 1, lane);
    __syncwarp();
    if (m <= 32) {
        const int key = lane < m ? seq[lane] : 0;
        const int slot = par_insert(key, lane < m, mask, lane);
        if (lane < m) t[slot] = (int16_t)key;
    } else if (m <= 32 * PAR_KPL && mask <= 127) {
        par_insert_multi(t, mask, seq, m, own, lane);
    } else if (lane == 0) {
        for (int i = 0; i < m; ++i) tab_insert_clean(t, mask, seq[i]);
    }
    __syncwarp();
}

// ---- warp-cooperative primitives ----------------------------------------------------
__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(FULL, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// ascending keys of (a [& | &~] f) -> out ; returns count
__device__ __noinline__ int bits_to_seq(const uint32_t* a, const uint32_t* f, bool want_in, int NW, int16_t* out, int lane) {
    int total = 0;
    for (int w0 = 0; w0 < NW; w0 += 32) {
        const int w = w0 + lane;
        uint32_t v = w < NW ? a[w] : 0u;
        if (f && w < NW) v = want_in ? (v & f[w]) : (v & ~f[w]);
        const int c = __popc(v);
        const int incl = warp_incl_scan(c, lane);
        int base = total + incl - c;
        while (v) { const int b = __ffs(v) - 1; out[base++] = (int16_t)(w * 32 + b); v &= v - 1; }
        total += __shfl_sync(FULL, incl, 31);
    }
    return total;
}

// live keys of explicit table in slot order, filtered -> out ; returns count
__device__ __noinline__ int tab_to_seq(const int16_t* t, int size, const uint32_t* f, bool want_in, int16_t* out, int lane) {
    int total = 0;
    for (int i0 = 0; i0 < size; i0 += 32) {
        const int i = i0 + lane;
        const int k = i < size ? t[i] : -1;
        bool ok = k >= 0;
        if (ok && f) ok = (((f[k >> 5] >> (k & 31)) & 1u) != 0) == want_in;
        const unsigned m = __ballot_sync(FULL, ok);
        if (ok) out[total + __popc(m & ((1u << lane) - 1))] = (int16_t)k;
        total += __popc(m);
    }
    return total;
}

// ordered live keys of a set (optionally filtered by bitset f)
__device__ __forceinline__ int set_to_seq(const SetRef& s, int K, int NWe, const uint32_t* f, bool want_in, int16_t* out, int lane) {
    const int size = s.mask() + 1;
    if (size >= K) return bits_to_seq(s.bits(), f, want_in, NWe, out, lane);
    return tab_to_seq(s.tab(), size, f, want_in, out, lane);
}

__device__ __forceinline__ int bits_count2(const uint32_t* a, const uint32_t* b, bool and_not, int NWe, int lane) {
    int c = 0;
    for (int w = lane; w < NWe; w += 32) c += __popc(and_not ? (a[w] & ~b[w]) : (a[w] & b[w]));
    return __reduce_add_sync(FULL, c);
}

// dst = fresh set holding (a op b); `seq` is the insertion order (needed only when the table
// is smaller than the node count); bits are always a op b.
__device__ __forceinline__ void fresh_from(const SetRef& dst, const uint32_t* a, const uint32_t* b, bool and_not, int n, int K,
                                           int NWe, const int16_t* seq, int16_t* tmp, uint32_t* own, int lane) {
    for (int w = lane; w < NWe; w += 32) dst.bits()[w] = and_not ? (a[w] & ~b[w]) : (a[w] & b[w]);
    const int size = growth_size(n);
    if (lane == 0) { dst.mask() = size - 1; dst.fill() = n; dst.used() = n; dst.finger() = 0; }
    if (size < K) tab_build_by_adds(dst.tab(), seq, n, tmp, own, lane);
    __syncwarp();
}

struct CliqueArgs {
    CliqueGeom g;
    int P;
    const int32_t* counts;     // [P] node count of each problem
    int Kmax;                  // row capacity of the per-problem arrays (== g.Kpad)
    const uint32_t* adjbits;   // [P][Kpad][NW]
    int16_t* adjseq;           // [P][Kpad][SEQCAP] slot-ordered neighbours of small-table nodes
    uint32_t* stack;           // [P][Kpad + 1][3 * SW + 1]  (subg | cand | ext | qn)
    int prune;
    const int32_t* max_size;   // [P] exact maximum clique size from k_maxclique (0 = unknown) or nullptr
    const uint32_t* vstar;     // [P][vstar_stride] vertices that can belong to a clique of that size (k_viable), or nullptr
    size_t vstar_stride;
    int adj_in_smem;           // adjacency rows staged in shared memory (row stride g.RS words)
    int adjseq_ready;          // adjseq was filled by k_adjseq (one warp per node) before this launch
    long long node_limit;
    // outputs
    uint8_t* mask;             // [P][Kpad]
    int32_t* n_inliers;        // [P]
    int32_t* nodes;            // [P] pops performed
    int32_t* status;           // [P]
    long long* n_yields;       // [P] (may be null)
    unsigned long long* order_hash;  // [P] (may be null) FNV-1a over (size, members...) of every yield
    long long* prof;           // [P][16] cycle counters per phase + event counts (null unless RF_CLIQUE_PROFILE is set)
};

__device__ __forceinline__ unsigned long long fnv_mix(unsigned long long hsh, int v) {
    for (int b = 0; b < 4; ++b) { hsh ^= (unsigned long long)((v >> (8 * b)) & 0xFF); hsh *= 1099511628211ULL; }
    return hsh;
}

// dst = a & adj[q], |result| = n > 0 already known.
// CPython set_intersection iterates the smaller operand (adj[q] on a tie) and adds the keys
// found in the other one, so the insertion order is the slot order of that operand.
__device__ __noinline__ void build_and_adj(const SetRef& dst, const SetRef& a, int n, int K, int NWe, const uint32_t* adjq,
                                           int degq, const int16_t* adjseq_q, int16_t* seq, int16_t* tmp, uint32_t* own, int lane) {
    if (growth_size(n) < K) {
        if (degq <= a.used()) {
            if (growth_size(degq) >= K) bits_to_seq(adjq, a.bits(), true, NWe, seq, lane);
            else {
                // neighbours of q in adj[q]'s slot order, keep those in a
                int m = 0;
                for (int i0 = 0; i0 < degq; i0 += 32) {
                    const int i = i0 + lane;
                    const int k = i < degq ? adjseq_q[i] : -1;
                    const bool ok = k >= 0 && ((a.bits()[k >> 5] >> (k & 31)) & 1u);
                    const unsigned bm = __ballot_sync(FULL, ok);
                    if (ok) seq[m + __popc(bm & ((1u << lane) - 1))] = (int16_t)k;
                    m += __popc(bm);
                }
            }
        } else {
            set_to_seq(a, K, NWe, adjq, true, seq, lane);
        }
        __syncwarp();
    }
    fresh_from(dst, a.bits(), adjq, false, n, K, NWe, seq, tmp, own, lane);
}

// dst = a - adj[u]   (CPython set_difference, both branches)
__device__ __noinline__ void set_sub_adj(const SetRef& dst, const SetRef& a, int K, int NWe, const uint32_t* adju, int degu,
                                         int16_t* seq, int16_t* tmp, uint32_t* own, int lane) {
    if ((a.used() >> 2) > degu) {
        // set_copy_and_difference: copy a (one up-front resize to used*2), then discard
        const int a_used = a.used(), a_fill = a.fill(), a_mask = a.mask();
        int newmask = 7;
        if (a_used * 5 >= 7 * 3) { int ns = 8; while (ns <= a_used * 2) ns <<= 1; newmask = ns - 1; }
        const bool explicit_tab = (newmask + 1) < K;
        int m = 0;
        if (explicit_tab) {
            if (newmask == a_mask && a_fill == a_used) {   // slot-for-slot copy
                for (int i = lane; i <= newmask; i += 32) dst.tab()[i] = a.tab()[i];
                __syncwarp();
            } else {
                m = set_to_seq(a, K, NWe, nullptr, true, seq, lane);
                __syncwarp();
                tab_reinsert(dst.tab(), newmask, seq, m, own, lane);
            }
        }
        const int removed = bits_count2(a.bits(), adju, false, NWe, lane);
        if (explicit_tab) {
            for (int i = lane; i <= newmask; i += 32) {
                const int k = dst.tab()[i];
                if (k >= 0 && ((adju[k >> 5] >> (k & 31)) & 1u)) dst.tab()[i] = DUMMY_SLOT;
            }
        }
        for (int w = lane; w < NWe; w += 32) dst.bits()[w] = a.bits()[w] & ~adju[w];
        __syncwarp();
        int fill = a_used, used = a_used - removed, mask = newmask;
        if ((fill - used) > mask / 4) {   // "more than 1/4 dummies": rebuild at used*4
            int ns = 8; while (ns <= used * 4) ns <<= 1;
            if (explicit_tab) {
                m = tab_to_seq(dst.tab(), mask + 1, nullptr, true, seq, lane);
                __syncwarp();
                if (ns < K) tab_reinsert(dst.tab(), ns - 1, seq, m, own, lane);
            } else if (ns < K) {
                m = bits_to_seq(dst.bits(), nullptr, true, NWe, seq, lane);
                __syncwarp();
                tab_reinsert(dst.tab(), ns - 1, seq, m, own, lane);
            }
            mask = ns - 1; fill = used;
        }
        if (lane == 0) { dst.mask() = mask; dst.fill() = fill; dst.used() = used; dst.finger() = 0; }
        __syncwarp();
        return;
    }
    const int n = bits_count2(a.bits(), adju, true, NWe, lane);
    if (n > 0 && growth_size(n) < K) { set_to_seq(a, K, NWe, adju, false, seq, lane); __syncwarp(); }
    fresh_from(dst, a.bits(), adju, true, n, K, NWe, seq, tmp, own, lane);
}

// ext.pop(): first live slot at or after finger (wrapping)
__device__ __noinline__ int set_pop(const SetRef& s, int K, int NWe, int lane) {
    const int mask = s.mask(), size = mask + 1;
    const int start = s.finger() & mask;
    int found = -1, slot = -1;
    if (size >= K) {
        // identity layout: slot == key.  next set bit >= start, else lowest set bit
        for (int pass = 0; pass < 2 && found < 0; ++pass) {
            const int from = pass == 0 ? start : 0;
            for (int w0 = (from >> 5); w0 < NWe && found < 0; w0 += 32) {
                const int w = w0 + lane;
                uint32_t v = w < NWe ? s.bits()[w] : 0u;
                if (w == (from >> 5)) v &= ~0u << (from & 31);
                const unsigned bm = __ballot_sync(FULL, v != 0u);
                if (bm) {
                    const int src = __ffs(bm) - 1;
                    const uint32_t vv = __shfl_sync(FULL, v, src);
                    found = (w0 + src) * 32 + (__ffs(vv) - 1);
                }
            }
        }
        slot = found;
    } else {
        const int16_t* t = s.tab();
        for (int pass = 0; pass < 2 && found < 0; ++pass) {
            const int from = pass == 0 ? start : 0;
            for (int i0 = from & ~31; i0 < size && found < 0; i0 += 32) {
                const int i = i0 + lane;
                const int k = (i < size && i >= from) ? t[i] : -1;
                const unsigned bm = __ballot_sync(FULL, k >= 0);
                if (bm) {
                    const int src = __ffs(bm) - 1;
                    found = __shfl_sync(FULL, k, src);
                    slot = i0 + src;
                }
            }
        }
        __syncwarp();                               // every lane has read the table before lane 0 rewrites a slot
        if (lane == 0) s.tab()[slot] = DUMMY_SLOT;
    }
    __syncwarp();                                   // ... and the bit set (racecheck: explicit read -> write order)
    if (lane == 0) {
        s.bits()[found >> 5] &= ~(1u << (found & 31));
        s.used() -= 1;
        s.finger() = slot + 1;
    }
    __syncwarp();
    return found;
}

// cand.remove(q)
__device__ __forceinline__ void set_remove(const SetRef& s, int K, int q, int lane) {
    const int size = s.mask() + 1;
    if (size < K) {
        int16_t* t = s.tab();
        for (int i = lane; i < size; i += 32) if (t[i] == q) t[i] = DUMMY_SLOT;
    }
    if (lane == 0) { s.bits()[q >> 5] &= ~(1u << (q & 31)); s.used() -= 1; }
    __syncwarp();
}

// words of a set that carry state: bitset + header (+ table when it is explicit)
__device__ __forceinline__ int set_live_words(const SetRef& s, int K) {
    const int size = s.mask() + 1;
    return s.NW + 4 + (size < K ? (size >> 1) : 0);
}
__device__ __forceinline__ void set_copy(uint32_t* dst, const SetRef& s, int K, int lane) {
    const int n = set_live_words(s, K);
    for (int i = lane; i < n; i += 32) dst[i] = s.w[i];
}

// Greedy sequential colouring of a vertex set (lane w holds word w of the bitset; NWe <= 32): an upper bound on the
// size of a clique inside the set.  Stops as soon as `need` colour classes are required (returns need): the caller
// only asks whether fewer than `need` vertices can be pairwise adjacent.
__device__ __forceinline__ int colour_bound(uint32_t Q, const uint32_t* __restrict__ adj, int RS, int NWe, int need, int lane) {
    int k = 0;
    while (__any_sync(FULL, Q != 0u)) {
        if (++k >= need) return need;
        uint32_t Qk = Q;
        for (;;) {
            const unsigned bm = __ballot_sync(FULL, Qk != 0u);
            if (!bm) break;
            const int src = __ffs(bm) - 1;
            const uint32_t w = __shfl_sync(FULL, Qk, src);
            const int b = __ffs(w) - 1;
            const int v = src * 32 + b;
            const uint32_t row = lane < NWe ? adj[(size_t)v * RS + lane] : 0u;
            Qk &= ~row;
            if (lane == src) { Qk &= ~(1u << b); Q &= ~(1u << b); }
        }
    }
    return k;
}

// V*: the vertices that can belong to a clique of the known maximum size M.  v stays only while its neighbourhood inside
// the current V needs at least M - 1 colours (so can hold M - 1 pairwise adjacent vertices); every member of a clique of
// size M survives every round (its M - 1 fellow members are in its neighbourhood and, by induction, still in V).  On the
// sparse graphs of freshly re-detected pairs (230 nodes, maximum clique 24) this leaves little more than the consistent
// set, and the order-exact walk stops descending into branches that only the contextual bound used to rule out — after
// having built their child sets.  One CTA of VI_WARPS warps per problem, double-buffered V (a round reads one copy and
// clears bits in the other).
#define VI_WARPS 16
__global__ void __launch_bounds__(32 * VI_WARPS) k_viable(CliqueGeom g, int P, const int32_t* __restrict__ counts,
                                                           const uint32_t* __restrict__ adjbits_all, const int32_t* __restrict__ max_size,
                                                           uint32_t* __restrict__ vstar_all, size_t vstar_stride) {
    extern __shared__ uint32_t sm_vi[];
    __shared__ uint32_t s_v[2][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int p = blockIdx.x;
    if (p >= P) return;
    const int K = counts[p];
    uint32_t* vs = vstar_all + (size_t)p * vstar_stride;
    const int NWe = (K + 31) >> 5;
    const int M = max_size[p];
    if (K <= 0 || M <= 1 || NWe > 32) {                      // unknown size: every vertex stays
        for (int w = tid; w < g.NW; w += 32 * VI_WARPS) vs[w] = 0xFFFFFFFFu;
        return;
    }
    const int RS = g.RS;
    const uint32_t* adj_gl = adjbits_all + (size_t)p * g.Kpad * g.NW;
    for (int t = tid; t < K * NWe; t += 32 * VI_WARPS) {
        const int r = t / NWe, w = t - r * NWe;
        sm_vi[r * RS + w] = adj_gl[(size_t)r * g.NW + w];
    }
    if (tid < 32) {
        const int lo = tid * 32;
        const uint32_t v = K >= lo + 32 ? ~0u : (K > lo ? ((1u << (K - lo)) - 1u) : 0u);
        s_v[0][tid] = v; s_v[1][tid] = v;
    }
    __syncthreads();
    const int need = M - 1;
    int cur = 0;
    for (int round = 0; round < 4; ++round) {
        const uint32_t* V = s_v[cur];
        uint32_t* Vn = s_v[cur ^ 1];
        for (int v = warp; v < K; v += VI_WARPS) {
            if (!((V[v >> 5] >> (v & 31)) & 1u)) continue;
            const uint32_t Q = lane < NWe ? (sm_vi[v * RS + lane] & V[lane]) : 0u;
            bool ok = __reduce_add_sync(FULL, __popc(Q)) >= need;
            if (ok) ok = colour_bound(Q, sm_vi, RS, NWe, need, lane) >= need;
            if (!ok && lane == 0) atomicAnd(&Vn[v >> 5], ~(1u << (v & 31)));
        }
        __syncthreads();
        // another round only pays while the set still shrinks fast (a round costs a colouring per vertex: ~0.1 ms on a dense
        // 200-node graph, where it removes little)
        int before = 0, after = 0;
        for (int w = 0; w < NWe; ++w) { before += __popc(s_v[cur][w]); after += __popc(s_v[cur ^ 1][w]); }
        __syncthreads();
        // the next round reads what this one wrote; bring the other copy up to date
        if (tid < 32) s_v[cur][tid] = s_v[cur ^ 1][tid];
        cur ^= 1;
        __syncthreads();
        if ((before - after) * 8 < after) break;
    }
    if (tid < g.NW) vs[tid] = tid < 32 ? s_v[cur][tid] : 0u;
}

// Slot orders of the small-table adjacency sets (adj[v] = {x for x in G[v] if x != v}: ascending adds into a table
// smaller than the node count), one WARP PER NODE.  Inside k_clique the same work is a sequential prologue of the one
// search warp (4.5 k cycles per node: 13 % of the search of a sparse 232-node graph, a third of the mean pair of the
// chained step).
#define ADJSEQ_WARPS 4
#define ADJSEQ_CTAS 8
__global__ void __launch_bounds__(32 * ADJSEQ_WARPS) k_adjseq(CliqueGeom g, int P, const int32_t* __restrict__ counts,
                                                               const uint32_t* __restrict__ adjbits_all, int16_t* __restrict__ adjseq_all) {
    extern __shared__ uint32_t sm_as[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int p = blockIdx.y;
    if (p >= P) return;
    const int K = counts[p];
    const int NWe = (K + 31) >> 5;
    const int own_words = 2 * g.NW > 128 ? 2 * g.NW : 128;
    const int per_warp = g.Kpad + g.TABN / 2 + own_words;        // seq + tmp (Kpad int16 each), table, arbitration scratch
    uint32_t* base = sm_as + (size_t)warp * per_warp;
    int16_t* seq = (int16_t*)base;
    int16_t* tmp = seq + g.Kpad;
    int16_t* tab = tmp + g.Kpad;
    uint32_t* own = base + g.Kpad + g.TABN / 2;
    // ADJSEQ_CTAS x ADJSEQ_WARPS warps per problem walk its nodes (a few thousand short CTAs per batch instead of one per
    // four nodes: most nodes of an inlier-dominated graph have large tables and nothing to do)
    for (int v = blockIdx.x * ADJSEQ_WARPS + warp; v < K; v += ADJSEQ_CTAS * ADJSEQ_WARPS) {
        const uint32_t* row = adjbits_all + ((size_t)p * g.Kpad + v) * g.NW;
        int c = 0;
        for (int w = lane; w < NWe; w += 32) c += __popc(row[w]);
        c = __reduce_add_sync(FULL, c);
        if (c <= 0 || growth_size(c) >= K) continue;
        const int m = bits_to_seq(row, nullptr, true, NWe, seq, lane);
        __syncwarp();
        tab_build_by_adds(tab, seq, m, tmp, own, lane);
        tab_to_seq(tab, growth_size(c), nullptr, true, adjseq_all + ((size_t)p * g.Kpad + v) * g.SEQCAP, lane);
        __syncwarp();
    }
}

#define PROF_MARK(slot) do { if (a.prof) { const long long _t = clock64(); pc[slot] += _t - tprev; tprev = _t; } } while (0)
__global__ void __launch_bounds__(32) k_clique(const CliqueArgs a) {
    long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#ifdef RF_CLIQUE_EVENTS
    long long ev[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // diagnostic event counts: node entries, chains, chain vertices, clique shortcuts, pivots, descents, bound prunes, leaf yields
#define EVT(i, n) (ev[i] += (n))
#else
#define EVT(i, n) ((void)0)
#endif
    long long tprev = clock64();
    extern __shared__ uint32_t sm[];
    const int lane = threadIdx.x;
    const int p = blockIdx.x;
    if (p >= a.P) return;
    const CliqueGeom g = a.g;
    const int K = a.counts[p];
    const int NW = g.NW, SW = g.SW;
    const int NWe = (K + 31) >> 5;       // words that can hold a set bit
    // shared layout: subg | cand | ext | ch_subg | ch_cand | seq[Kpad] | tmp[Kpad] | Q[Kpad] | bestQ[Kpad] | deg[Kpad] | adj rows
    SetRef subg{sm, NW}, cand{sm + SW, NW}, ext{sm + 2 * SW, NW}, chs{sm + 3 * SW, NW}, chc{sm + 4 * SW, NW};
    int16_t* seq = (int16_t*)(sm + 5 * SW);
    int16_t* tmp = seq + g.Kpad;
    int16_t* Q = tmp + g.Kpad;
    int16_t* bestQ = Q + g.Kpad;
    int16_t* deg = bestQ + g.Kpad;
    uint32_t* own = (uint32_t*)(deg + g.Kpad);     // [max(128, 2 NW)] slot arbitration / bitset scratch
    uint32_t* vst = own + (2 * NW > 128 ? 2 * NW : 128);   // [NW] viable vertices (all ones unless k_viable ran)
    uint32_t* adj_sm = vst + NW;
    const uint32_t* adj_gl = a.adjbits + (size_t)p * g.Kpad * NW;
    const uint32_t* adjbits = a.adj_in_smem ? adj_sm : adj_gl;
    const int RS = a.adj_in_smem ? g.RS : NW;       // row stride in words
    int16_t* adjseq = a.adjseq + (size_t)p * g.Kpad * g.SEQCAP;
    const int FS = 3 * SW + 1;           // search frame: three sets + the depth of Q
    uint32_t* stack = a.stack + (size_t)p * (g.Kpad + 1) * FS;
    uint8_t* outmask = a.mask + (size_t)p * g.Kpad;

    for (int i = lane; i < g.Kpad; i += 32) outmask[i] = 0;
    if (K <= 0) {
        if (lane == 0) { a.n_inliers[p] = 0; a.nodes[p] = 0; a.status[p] = RF_OK; if (a.n_yields) a.n_yields[p] = 0; if (a.order_hash) a.order_hash[p] = 14695981039346656037ULL; }
        return;
    }
    if (a.adj_in_smem) {
        for (int t = lane; t < K * NWe; t += 32) {
            const int r = t / NWe, w = t - r * NWe;
            adj_sm[r * RS + w] = adj_gl[(size_t)r * NW + w];
        }
        __syncwarp();
    }
    PROF_MARK(0);   // staging
    // ---- degrees (one lane per node) and slot orders of small-table adjacency sets -------
    for (int u0 = 0; u0 < K; u0 += 32) {
        const int u = u0 + lane;
        int c = 0;
        if (u < K) {
            const uint32_t* row = adjbits + (size_t)u * RS;
            for (int w = 0; w < NWe; ++w) c += __popc(row[w]);
            deg[u] = (int16_t)c;
        }
        unsigned need = a.adjseq_ready ? 0u : __ballot_sync(FULL, u < K && c > 0 && growth_size(c) < K);
        while (need) {
            const int src = __ffs(need) - 1;
            need &= need - 1;
            const int v = u0 + src;
            const int cv = __shfl_sync(FULL, c, src);
            // adj[v] = {x for x in G[v] if x != v}: ascending adds into a small table
            const int m = bits_to_seq(adjbits + (size_t)v * RS, nullptr, true, NWe, seq, lane);
            __syncwarp();
            tab_build_by_adds(chs.tab(), seq, m, tmp, own, lane);
            tab_to_seq(chs.tab(), growth_size(cv), nullptr, true, adjseq + (size_t)v * g.SEQCAP, lane);
            __syncwarp();
        }
    }
    __syncwarp();
    PROF_MARK(1);   // degrees + adjacency slot orders
    // ---- root: cand = set(G) ; subg = cand.copy() --------------------------------------
    for (int w = lane; w < NW; w += 32) {
        const int lo = w * 32;
        uint32_t v = K >= lo + 32 ? ~0u : (K > lo ? ((1u << (K - lo)) - 1u) : 0u);
        cand.bits()[w] = v; subg.bits()[w] = v;
    }
    if (lane == 0) {
        const int size = growth_size(K);
        cand.mask() = size - 1; cand.fill() = K; cand.used() = K; cand.finger() = 0;
        subg.mask() = size - 1; subg.fill() = K; subg.used() = K; subg.finger() = 0;
    }
    __syncwarp();
    int qn = 1, sp = 0, best = 0, nbest = 0;    // nbest: members currently recorded in bestQ
    long long pops = 0, ny = 0;
    unsigned long long hsh = 14695981039346656037ULL;
    int status = RF_OK;
    // Production mode (no order hash requested) adds three ORDER-SAFE accelerations; the test hook
    // with the hash walks every level exactly like networkx does.
    const bool fast = a.order_hash == nullptr;

    // (1) Lower bound on the answer: a clique found by min-degree peeling has size L <= M (the
    // maximum), and the reference's answer is the first clique of size M.  Starting the
    // "strictly larger" test at L - 1 only discards cliques smaller than L, none of which can be
    // the answer, so children that cannot reach L are skipped from the very first level.
    const int Mknown = (fast && a.prune && a.max_size && NWe <= 32) ? a.max_size[p] : 0;   // (1b) below: it replaces this bound
    if (fast && a.prune && K > 1 && Mknown <= 0) {
        uint32_t* S = own;                         // NWe words (scratch is free until the first table build)
        int16_t* dS = tmp;                         // degree within S
        for (int w = lane; w < NWe; w += 32) S[w] = cand.bits()[w];
        for (int v = lane; v < K; v += 32) dS[v] = deg[v];
        __syncwarp();
        int ns = K;
        for (;;) {
            unsigned mk = 0xFFFFFFFFu;
            for (int v = lane; v < K; v += 32)
                if ((S[v >> 5] >> (v & 31)) & 1u) mk = min(mk, ((unsigned)dS[v] << 16) | (unsigned)v);
            mk = __reduce_min_sync(FULL, mk);
            if ((int)(mk >> 16) >= ns - 1) break;  // every member is adjacent to all the others
            const int r = (int)(mk & 0xFFFF);
            __syncwarp();
            if (lane == 0) S[r >> 5] &= ~(1u << (r & 31));
            --ns;
            const uint32_t* rowr = adjbits + (size_t)r * RS;
            for (int v = lane; v < K; v += 32)
                if ((rowr[v >> 5] >> (v & 31)) & 1u) dS[v] -= 1;
            __syncwarp();
        }
        best = ns - 1;
    }
    // (1b) The exact maximum M (k_maxclique: order-free colouring branch-and-bound over several warps): the answer is the
    // FIRST clique of size M in networkx order, so only subtrees that can still hold M vertices are walked, with the
    // colouring bound instead of |cand|, and the search stops at the first clique of size M.
    const int Mmax = Mknown;
    if (Mmax > 0) best = Mmax - 1;
    // (1c) vertices outside V* (k_viable) are in no clique of size M: a branch on one of them is never descended, and
    // the colouring bound of a branch only counts candidates inside V*
    const bool use_vst = Mmax > 0 && a.vstar != nullptr;
    for (int w = lane; w < NW; w += 32) vst[w] = use_vst ? a.vstar[(size_t)p * a.vstar_stride + w] : 0xFFFFFFFFu;
    __syncwarp();
    PROF_MARK(2);   // greedy bound

    auto enter_node = [&]() {
        for (;;) {
            EVT(0, 1);
            // u = max(subg, key=lambda u: len(cand & adj[u])) — first maximum in subg's iteration order.
            // Iteration order = slot order; a lane scores the key of "its" slots, ties go to the lowest slot.
            const int size = subg.mask() + 1;
            const bool ident = size >= K;          // identity layout: slot == key
            const int nslots = ident ? K : size;
            const int ncand = cand.used();
            unsigned bestkey = 0;
            int n_univ = 0, max_out = -1;
            uint32_t* D = own;                     // universal candidates (ident layout only): NWe words
            for (int i0 = 0; i0 < nslots; i0 += 32) {
                const int i = i0 + lane;
                int u = -1;
                if (i < nslots) {
                    if (ident) u = ((subg.bits()[i >> 5] >> (i & 31)) & 1u) ? i : -1;
                    else u = subg.tab()[i];
                }
                bool univ = false;
                if (u >= 0) {
                    const uint32_t* row = adjbits + (size_t)u * RS;
                    int c = 0;
                    for (int w = 0; w < NWe; ++w) c += __popc(cand.bits()[w] & row[w]);
                    bestkey = max(bestkey, ((unsigned)c << 16) | (unsigned)(0xFFFF - i));
                    if ((cand.bits()[u >> 5] >> (u & 31)) & 1u) univ = c == ncand - 1;
                    else max_out = max(max_out, c);
                }
                if (fast) {
                    const unsigned bm = __ballot_sync(FULL, univ);
                    n_univ += __popc(bm);
                    __syncwarp();
                    if (ident && lane == 0) D[i0 >> 5] = bm;
                }
            }
            bestkey = __reduce_max_sync(FULL, bestkey);
            if (fast) {
                max_out = __reduce_max_sync(FULL, max_out);
                __syncwarp();
                // (2) cand is a clique: the subtree yields at most ONE maximal clique, Q + cand — and only
                // if no excluded vertex (subg \ cand) is adjacent to all of cand — and leaves the parent's
                // sets untouched, so its |cand| levels need not be walked.
                if (n_univ == ncand) {
                    EVT(3, 1);
                    if (max_out < ncand) {
                        ++ny;
                        const int csize = qn - 1 + ncand;
                        if (csize > best) {                      // outlierRejection.py:73 strict '>'
                            best = csize; nbest = csize;
                            for (int i = lane; i < qn - 1; i += 32) bestQ[i] = Q[i];
                            __syncwarp();
                            bits_to_seq(cand.bits(), nullptr, true, NWe, bestQ + (qn - 1), lane);
                            __syncwarp();
                        }
                    }
                    if (lane == 0) ext.used() = 0;               // nothing left to expand at this node
                    __syncwarp();
                    PROF_MARK(3);
                    return;
                }
                // (3) chain of universal candidates.  Let D = candidates adjacent to every other candidate.
                // If no excluded vertex reaches that score, the pivot is a member of D, ext = {pivot}, the
                // only child is (Q + pivot, subg & N(pivot), cand - pivot), and there D - pivot is again the
                // set of universal candidates: |D| levels without any branching or yield.  While the sets
                // stay in identity layout (table >= node count) their state after the chain is independent
                // of the order in which D was consumed, so the chain is taken in one step.
                if (n_univ > 0 && max_out < ncand - 1 && NWe <= 32) {
                    if (a.prune && qn - 1 + ncand <= best) {     // the single child cannot beat the best
                        if (lane == 0) ext.used() = 0;
                        __syncwarp();
                        return;
                    }
                    // D by key, word `lane` in a register (the table builds below use `own` as scratch)
                    uint32_t dreg = 0u;
                    if (ident) dreg = lane < NWe ? D[lane] : 0u;
                    else {
                        for (int i0 = 0; i0 < K; i0 += 32) {
                            const int v = i0 + lane;
                            bool uv = false;
                            if (v < K && ((cand.bits()[v >> 5] >> (v & 31)) & 1u)) {
                                const uint32_t* row = adjbits + (size_t)v * RS;
                                int c = 0;
                                for (int w = 0; w < NWe; ++w) c += __popc(cand.bits()[w] & row[w]);
                                uv = c == ncand - 1;
                            }
                            const unsigned bm = __ballot_sync(FULL, uv);
                            if (lane == (i0 >> 5)) dreg = bm;
                        }
                    }
                    int left = n_univ, ncur = ncand;
                    // (3a) identity layout: the first m members of D (networkx consumes D in ascending order there) in ONE
                    // step, m the largest count that leaves both sets in identity layout (|cand| - m elements still grow
                    // a table of at least K slots)
                    if (ident) {
                        const int nmin = K <= 8 ? 0 : K <= 32 ? 5 : K <= 128 ? 19 : K <= 512 ? 77 : K <= 2048 ? 307 : 1229;
                        const int m = min(left, ncur - nmin);
                        if (m >= 1) {
                            uint32_t dsel = dreg;
                            if (m < left) {                      // keep the m lowest members
                                const int c = __popc(dsel), incl = warp_incl_scan(c, lane), excl = incl - c;
                                if (excl >= m) dsel = 0u;
                                else while (excl + __popc(dsel) > m) dsel &= ~(0x80000000u >> __clz(dsel));
                            }
                            __syncwarp();
                            if (lane < NWe) D[lane] = dsel;
                            __syncwarp();
                            const int nc2 = ncur - m;
                            // subg'' = members of subg adjacent to all of the m
                            uint32_t* S2 = own + NW;
                            int ns2 = 0;
                            for (int i0 = 0; i0 < K; i0 += 32) {
                                const int v = i0 + lane;
                                bool keep = false;
                                if (v < K && ((subg.bits()[v >> 5] >> (v & 31)) & 1u)) {
                                    const uint32_t* row = adjbits + (size_t)v * RS;
                                    int c = 0;
                                    for (int w = 0; w < NWe; ++w) c += __popc(D[w] & row[w]);
                                    keep = c == m;
                                }
                                const unsigned bm = __ballot_sync(FULL, keep);
                                ns2 += __popc(bm);
                                if (lane == 0) S2[i0 >> 5] = bm;
                            }
                            __syncwarp();
                            if (growth_size(nc2) >= K && growth_size(ns2) >= K) {
                                EVT(1, 1); EVT(2, m);
                                bits_to_seq(D, nullptr, true, NWe, Q + (qn - 1), lane);
                                qn += m;
                                pops += m;
                                for (int w = lane; w < NWe; w += 32) { cand.bits()[w] &= ~D[w]; subg.bits()[w] = S2[w]; }
                                if (lane == 0) {
                                    Q[qn - 1] = -1;
                                    cand.mask() = growth_size(nc2) - 1; cand.fill() = nc2; cand.used() = nc2; cand.finger() = 0;
                                    subg.mask() = growth_size(ns2) - 1; subg.fill() = ns2; subg.used() = ns2; subg.finger() = 0;
                                }
                                __syncwarp();
                                dreg &= ~dsel; left -= m; ncur -= m;
                            }
                        }
                    }
                    // (3b) the remaining members of D, one level each, WITHOUT re-scoring: the pivot of every level of the
                    // chain is the first member of D in subg's iteration order (all of D tie at the maximum score and no
                    // excluded vertex reaches it), ext = {pivot}, and D - pivot is the child's set of universal candidates.
                    // Only the two child sets are built (their slot layout is what the next level iterates); scoring,
                    // ext = cand - adj[u], ext.pop() and the frame push of the general path (two thirds of a level) are
                    // skipped.  Stops when cand has become a clique: shortcut (2) finishes that on the next pass.
                    while (left > 0 && left < ncur) {
                        const int ssz = subg.mask() + 1;
                        int u = -1;
                        if (ssz >= K) {
                            const unsigned bm = __ballot_sync(FULL, dreg != 0u);
                            const int src = __ffs(bm) - 1;
                            const uint32_t w = __shfl_sync(FULL, dreg, src);
                            u = src * 32 + (__ffs(w) - 1);
                        } else {
                            const int16_t* t = subg.tab();
                            for (int i0 = 0; i0 < ssz && u < 0; i0 += 32) {
                                const int i = i0 + lane;
                                const int k = i < ssz ? t[i] : -1;
                                const uint32_t dw = __shfl_sync(FULL, dreg, (k >= 0 ? k : 0) >> 5);
                                const bool hit = k >= 0 && ((dw >> (k & 31)) & 1u);
                                const unsigned bm = __ballot_sync(FULL, hit);
                                if (bm) u = __shfl_sync(FULL, k, __ffs(bm) - 1);
                            }
                        }
                        EVT(1, 1); EVT(2, 1);
                        ++pops;
                        __syncwarp();
                        set_remove(cand, K, u, lane);
                        if (lane == 0) Q[qn - 1] = (int16_t)u;
                        const uint32_t* adju = adjbits + (size_t)u * RS;
                        const int nsub = bits_count2(subg.bits(), adju, false, NWe, lane);
                        const int nc = bits_count2(cand.bits(), adju, false, NWe, lane);
                        const int16_t* adjseq_u = adjseq + (size_t)u * g.SEQCAP;
                        build_and_adj(chs, subg, nsub, K, NWe, adju, deg[u], adjseq_u, seq, tmp, own, lane);
                        build_and_adj(chc, cand, nc, K, NWe, adju, deg[u], adjseq_u, seq, tmp, own, lane);
                        if (lane == 0) Q[qn] = -1;
                        __syncwarp();
                        ++qn;
                        set_copy(subg.w, chs, K, lane);
                        set_copy(cand.w, chc, K, lane);
                        __syncwarp();
                        if (lane == (u >> 5)) dreg &= ~(1u << (u & 31));
                        --left; --ncur;
                    }
                    continue;
                }
                if (ident && n_univ > 0 && max_out < ncand - 1) {   // NWe > 32: all of D or nothing
                    if (a.prune && qn - 1 + ncand <= best) {     // the single child cannot beat the best
                        if (lane == 0) ext.used() = 0;
                        __syncwarp();
                        return;
                    }
                    const int nc2 = ncand - n_univ;
                    // subg'' = members of subg adjacent to all of D
                    uint32_t* S2 = own + NW;
                    int ns2 = 0;
                    for (int i0 = 0; i0 < K; i0 += 32) {
                        const int v = i0 + lane;
                        bool keep = false;
                        if (v < K && ((subg.bits()[v >> 5] >> (v & 31)) & 1u)) {
                            const uint32_t* row = adjbits + (size_t)v * RS;
                            int c = 0;
                            for (int w = 0; w < NWe; ++w) c += __popc(D[w] & row[w]);
                            keep = c == n_univ;
                        }
                        const unsigned bm = __ballot_sync(FULL, keep);
                        ns2 += __popc(bm);
                        if (lane == 0) S2[i0 >> 5] = bm;
                    }
                    __syncwarp();
                    if (growth_size(nc2) >= K && growth_size(ns2) >= K) {
                        EVT(1, 1); EVT(2, n_univ);
                        bits_to_seq(D, nullptr, true, NWe, Q + (qn - 1), lane);
                        qn += n_univ;
                        pops += n_univ;
                        for (int w = lane; w < NWe; w += 32) { cand.bits()[w] &= ~D[w]; subg.bits()[w] = S2[w]; }
                        if (lane == 0) {
                            Q[qn - 1] = -1;
                            cand.mask() = growth_size(nc2) - 1; cand.fill() = nc2; cand.used() = nc2; cand.finger() = 0;
                            subg.mask() = growth_size(ns2) - 1; subg.fill() = ns2; subg.used() = ns2; subg.finger() = 0;
                        }
                        __syncwarp();
                        continue;
                    }
                }
            }
            const int slot = 0xFFFF - (int)(bestkey & 0xFFFF);
            const int u = ident ? slot : subg.tab()[slot];
            EVT(4, 1);
            PROF_MARK(3);   // pivot scoring / shortcut / chain
            set_sub_adj(ext, cand, K, NWe, adjbits + (size_t)u * RS, deg[u], seq, tmp, own, lane);
            PROF_MARK(4);   // ext = cand - adj[u]
            return;
        }
    };
    enter_node();

    for (;;) {
        if (Mmax > 0 && nbest >= Mmax) break;
        if (ext.used() > 0) {
            if (pops >= a.node_limit) { status = RF_E_WORKLIMIT; break; }
            ++pops;
            const int q = set_pop(ext, K, NWe, lane);
            set_remove(cand, K, q, lane);
            if (lane == 0) Q[qn - 1] = (int16_t)q;
            const uint32_t* adjq = adjbits + (size_t)q * RS;
            const int nsub = bits_count2(subg.bits(), adjq, false, NWe, lane);
            PROF_MARK(5);   // pop + remove + count
            if (nsub == 0) {
                ++ny; EVT(7, 1);
                __syncwarp();
                if (a.order_hash) { hsh = fnv_mix(hsh, qn); for (int i = 0; i < qn; ++i) hsh = fnv_mix(hsh, Q[i]); }
                if (qn > best) {   // outlierRejection.py:73 strict '>'
                    best = qn; nbest = qn;
                    for (int i = lane; i < qn; i += 32) bestQ[i] = Q[i];
                    __syncwarp();
                    if (Mmax > 0 && best >= Mmax) break;      // the first maximum clique: nothing later can be strictly larger
                }
            } else {
                const int ncand = bits_count2(cand.bits(), adjq, false, NWe, lane);
                // a child is descended only if it can still beat the best clique so far; the parent's
                // sets do not depend on that decision, so the order of later yields is unchanged
                bool descend = ncand > 0 && !(a.prune && qn + ncand <= best);
                if (descend && Mmax > 0) {
                    if (!((vst[q >> 5] >> (q & 31)) & 1u)) descend = false;
                    else {
                        const uint32_t cw = lane < NWe ? (cand.bits()[lane] & adjq[lane] & vst[lane]) : 0u;
                        descend = qn + __reduce_add_sync(FULL, __popc(cw)) > best &&
                                  qn + colour_bound(cw, adjbits, RS, NWe, best - qn + 1, lane) > best;
                    }
                }
                if (!descend) EVT(6, 1);
                if (descend) {
                    EVT(5, 1);
                    const int degq = deg[q];
                    const int16_t* adjseq_q = adjseq + (size_t)q * g.SEQCAP;
                    build_and_adj(chs, subg, nsub, K, NWe, adjq, degq, adjseq_q, seq, tmp, own, lane);
                    build_and_adj(chc, cand, ncand, K, NWe, adjq, degq, adjseq_q, seq, tmp, own, lane);
                    PROF_MARK(6);   // child sets
                    // push the parent frame (live words only) and the depth of Q to return to — unless q was
                    // the parent's last child: an exhausted frame would only be popped and discarded
                    if (ext.used() > 0) {
                        uint32_t* fr = stack + (size_t)sp * FS;
                        set_copy(fr, subg, K, lane); set_copy(fr + SW, cand, K, lane); set_copy(fr + 2 * SW, ext, K, lane);
                        if (lane == 0) fr[3 * SW] = (uint32_t)qn;
                        ++sp;
                    }
                    if (lane == 0) Q[qn] = -1;
                    __syncwarp();
                    ++qn;
                    set_copy(subg.w, chs, K, lane);
                    set_copy(cand.w, chc, K, lane);
                    __syncwarp();
                    PROF_MARK(7);   // frame push
                    enter_node();
                    if (Mmax > 0 && nbest >= Mmax) break;     // found by the cand-is-a-clique shortcut
                }
            }
        } else {
            if (sp == 0) break;
            --sp;
            uint32_t* fr = stack + (size_t)sp * FS;
            __syncwarp();
            // header first (it says how many words are live), then the rest
            SetRef f0{fr, NW}, f1{fr + SW, NW}, f2{fr + 2 * SW, NW};
            set_copy(subg.w, f0, K, lane); set_copy(cand.w, f1, K, lane); set_copy(ext.w, f2, K, lane);
            qn = (int)fr[3 * SW];
            __syncwarp();
            PROF_MARK(7);   // frame pop
        }
    }
    __syncwarp();
    for (int i = lane; i < nbest; i += 32) outmask[bestQ[i]] = 1;
    if (lane == 0) {
        a.n_inliers[p] = nbest;
        a.nodes[p] = (int32_t)(pops > 0x7fffffff ? 0x7fffffff : pops);
        a.status[p] = status;
        if (a.n_yields) a.n_yields[p] = ny;
        if (a.order_hash) a.order_hash[p] = hsh;
        if (a.prof) for (int k = 0; k < 8; ++k) {
            a.prof[(size_t)p * 16 + k] = pc[k];
#ifdef RF_CLIQUE_EVENTS
            a.prof[(size_t)p * 16 + 8 + k] = ev[k];
#else
            a.prof[(size_t)p * 16 + 8 + k] = 0;
#endif
        }
    }
}


// ------------------------------------------------------------------------------------
// Exact maximum clique SIZE, order-free: bitset branch and bound with greedy colouring bounds
// (Tomita & Seki's MCQ on bitsets, as in San Segundo's BBMC).  It only has to answer "how large", so it is
// free of the set-order emulation above and can use several warps per problem: the root level is coloured
// once, its branches (vertex i of the colouring order + its earlier neighbours) are handed out to the warps
// from the most promising end, and the incumbent size is shared through shared memory.
// Vertices are relabelled by descending degree first (better colourings, smaller trees).
// One CTA of MC_WARPS warps per problem; K <= 1024 (32 words per bitset: one word per lane).
// ------------------------------------------------------------------------------------
#define MC_WARPS 4
#define MC_ARENA 4096          // colouring entries per warp (vertex | colour << 16); overflow -> size unknown (0)

struct MaxCliqueArgs {
    int P, Kpad, NW;
    const int32_t* counts;
    const uint32_t* adjbits;   // [P][Kpad][NW]
    int32_t* max_size;         // [P]
    uint32_t* levels;          // [P][MC_WARPS][Kpad][NW]   candidate set of every DFS level
    uint32_t* arena;           // [P][MC_WARPS][MC_ARENA]
    int32_t* lvl_meta;         // [P][MC_WARPS][Kpad][2]     (arena base, entries left)
};

// colour the set held one word per lane; entries with colour >= kmin are appended to out[] (vertex | colour << 16) in
// colouring order; returns the number of entries written (or -1 if more than cap), *ncol = colour classes used
__device__ __forceinline__ int mc_colour(uint32_t Q, const uint32_t* __restrict__ adj, int RSa, int NWe, int kmin, uint32_t* out,
                                         int cap, int lane, int* ncol) {
    int k = 0, idx = 0;
    while (__any_sync(FULL, Q != 0u)) {
        ++k;
        uint32_t Qk = Q;
        for (;;) {
            const unsigned bm = __ballot_sync(FULL, Qk != 0u);
            if (!bm) break;
            const int src = __ffs(bm) - 1;
            const uint32_t w = __shfl_sync(FULL, Qk, src);
            const int b = __ffs(w) - 1;
            const int v = src * 32 + b;
            const uint32_t row = lane < NWe ? adj[(size_t)v * RSa + lane] : 0u;
            Qk &= ~row;
            if (lane == src) { Qk &= ~(1u << b); Q &= ~(1u << b); }
            if (k >= kmin) {
                if (idx >= cap) { *ncol = k; return -1; }
                if (lane == 0) out[idx] = (uint32_t)v | ((uint32_t)k << 16);
                ++idx;
            }
        }
    }
    *ncol = k;
    return idx;
}

__global__ void __launch_bounds__(32 * MC_WARPS) k_maxclique(const MaxCliqueArgs a) {
    extern __shared__ uint32_t msm[];
    __shared__ int s_best, s_next, s_fail, s_nroot;
    const int p = blockIdx.x;
    const int K = a.counts[p];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (K <= 1 || K > 1024) { if (tid == 0) a.max_size[p] = K <= 1 ? max(K, 0) : 0; return; }
    const int NWe = (K + 31) >> 5, RSa = NWe | 1;
    uint32_t* adj = msm;                                   // [K][RSa] relabelled adjacency
    uint16_t* rank = (uint16_t*)(adj + (size_t)K * RSa);   // [K] new label of every vertex
    uint16_t* deg = rank + K;
    uint32_t* root = (uint32_t*)(deg + K);                 // [K] root colouring entries (rank + deg = K words: aligned)
    const uint32_t* g = a.adjbits + (size_t)p * a.Kpad * a.NW;
    if (tid == 0) { s_best = 0; s_fail = 0; }
    for (int v = tid; v < K; v += blockDim.x) {
        int c = 0;
        for (int w = 0; w < NWe; ++w) c += __popc(g[(size_t)v * a.NW + w]);
        deg[v] = (uint16_t)c;
    }
    for (int i = tid; i < K * RSa; i += blockDim.x) adj[i] = 0u;
    __syncthreads();
    for (int v = tid; v < K; v += blockDim.x) {            // rank by (degree desc, index asc)
        const int dv = deg[v];
        int r = 0;
        for (int u = 0; u < K; ++u) { const int du = deg[u]; r += (du > dv) || (du == dv && u < v); }
        rank[v] = (uint16_t)r;
    }
    __syncthreads();
    for (int v = tid; v < K; v += blockDim.x) {            // row rank[v] is written by this thread only
        uint32_t* row = adj + (size_t)rank[v] * RSa;
        for (int w = 0; w < NWe; ++w) {
            uint32_t bits = g[(size_t)v * a.NW + w];
            while (bits) { const int b = __ffs(bits) - 1; bits &= bits - 1; const int u = rank[w * 32 + b]; row[u >> 5] |= 1u << (u & 31); }
        }
    }
    __syncthreads();
    // ---- incumbent: greedy clique (always take the candidate with most neighbours among the candidates) ----
    if (wid == 0) {
        uint32_t Pw = lane < NWe ? ((K >= lane * 32 + 32) ? ~0u : (K > lane * 32 ? ((1u << (K - lane * 32)) - 1u) : 0u)) : 0u;
        int size = 0;
        while (__any_sync(FULL, Pw != 0u)) {
            unsigned bestkey = 0;
            for (int v0 = 0; v0 < K; v0 += 32) {
                const int v = v0 + lane;
                const uint32_t pv = __shfl_sync(FULL, Pw, v0 >> 5);
                const bool valid = v < K && ((pv >> (v & 31)) & 1u);
                int c = 0;
                for (int w = 0; w < NWe; ++w) {
                    const uint32_t pw = __shfl_sync(FULL, Pw, w);
                    if (valid) c += __popc(pw & adj[(size_t)v * RSa + w]);
                }
                if (valid) bestkey = max(bestkey, ((unsigned)c << 16) | (unsigned)(0xFFFF - v));
            }
            bestkey = __reduce_max_sync(FULL, bestkey);
            const int v = 0xFFFF - (int)(bestkey & 0xFFFF);
            ++size;
            Pw &= lane < NWe ? adj[(size_t)v * RSa + lane] : 0u;
        }
        // ---- root colouring ----
        uint32_t Pall = lane < NWe ? ((K >= lane * 32 + 32) ? ~0u : (K > lane * 32 ? ((1u << (K - lane * 32)) - 1u) : 0u)) : 0u;
        int ncol = 0;
        const int n = mc_colour(Pall, adj, RSa, NWe, 1, root, K, lane, &ncol);
        if (lane == 0) { s_best = size; s_nroot = n; s_next = n - 1; if (ncol <= size) s_next = -1; }   // colours == incumbent: optimal
        // Inlier-dominated graphs (a greedy clique of at least half the vertices): the order-exact walk with its
        // peeling bound is already short there (chains of universal candidates are taken in one step), while proving
        // the exact maximum costs a colouring per level of a ~100-deep tree and the colour bound a colouring per pop.
        // Report "unknown" and let k_clique run on its own bound.
        if (lane == 0 && 2 * size >= K) { s_fail = 1; s_next = -1; }
    }
    __syncthreads();
    const int nroot = s_nroot;
    uint32_t* levels = a.levels + ((size_t)p * MC_WARPS + wid) * (size_t)a.Kpad * a.NW;
    uint32_t* arena = a.arena + ((size_t)p * MC_WARPS + wid) * MC_ARENA;
    int32_t* meta = a.lvl_meta + ((size_t)p * MC_WARPS + wid) * (size_t)a.Kpad * 2;
    volatile int* vbest = &s_best;
    // position of every (relabelled) vertex in the root order, for the prefix sets of the root branches
    uint16_t* rpos = deg;                                  // deg is no longer needed
    for (int i = tid; i < nroot; i += blockDim.x) rpos[root[i] & 0xFFFF] = (uint16_t)i;
    __syncthreads();
    for (;;) {
        int i = 0;
        if (lane == 0) i = atomicSub(&s_next, 1);
        i = __shfl_sync(FULL, i, 0);
        if (i < 0 || s_fail) break;
        const uint32_t e = root[i];
        const int v = (int)(e & 0xFFFF), c = (int)(e >> 16);
        if (c <= *vbest) break;                           // every remaining root branch has a colour <= c
        // P0 = N(v) restricted to the vertices before position i of the root order
        uint32_t P0 = lane < NWe ? adj[(size_t)v * RSa + lane] : 0u;
        { uint32_t t = P0; while (t) { const int b = __ffs(t) - 1; t &= t - 1; if (rpos[lane * 32 + b] >= i) P0 &= ~(1u << b); } }
        if (!__any_sync(FULL, P0 != 0u)) { if (lane == 0 && 1 > *vbest) atomicMax(&s_best, 1); continue; }
        int d = 0;                                         // DFS level; the clique built so far has d + 1 vertices
        if (lane < NWe) levels[lane] = P0;
        bool enter = true;
        int top = 0;                                       // arena entries in use
        bool failed = false;
        for (;;) {
            uint32_t* Lp = levels + (size_t)d * a.NW;
            if (enter) {
                const uint32_t Pw = lane < NWe ? Lp[lane] : 0u;
                int ncol = 0;
                const int kmin = max(*vbest - (d + 1) + 1, 1);
                const int n = mc_colour(Pw, adj, RSa, NWe, kmin, arena + top, MC_ARENA - top, lane, &ncol);
                if (n < 0) { failed = true; break; }
                if (lane == 0) { meta[2 * d] = top; meta[2 * d + 1] = n; }
                top += n;
                __syncwarp();
                enter = false;
            }
            const int base = meta[2 * d];
            int left = meta[2 * d + 1];
            bool pop = left <= 0;
            uint32_t en = 0;
            if (!pop) { en = arena[base + left - 1]; pop = (d + 1) + (int)(en >> 16) <= *vbest; }
            if (pop) {
                if (d == 0) break;
                top = base;
                --d;
                continue;
            }
            const int u = (int)(en & 0xFFFF);
            __syncwarp();
            if (lane == 0) meta[2 * d + 1] = left - 1;
            uint32_t Pw = lane < NWe ? Lp[lane] : 0u;
            const uint32_t row = lane < NWe ? adj[(size_t)u * RSa + lane] : 0u;
            const uint32_t newP = Pw & row;
            if (lane == (u >> 5)) Lp[lane] = Pw & ~(1u << (u & 31));
            __syncwarp();
            const int cnt = __reduce_add_sync(FULL, __popc(newP));
            const int size = d + 2;                        // clique with u
            if (cnt == 0) { if (lane == 0 && size > *vbest) atomicMax(&s_best, size); continue; }
            if (size + cnt <= *vbest) continue;
            ++d;
            if (lane < NWe) levels[(size_t)d * a.NW + lane] = newP;
            __syncwarp();
            enter = true;
        }
        if (failed) { if (lane == 0) s_fail = 1; break; }
    }
    __syncthreads();
    if (tid == 0) a.max_size[p] = s_fail ? 0 : s_best;
}

// ------------------------------------------------------------------------------------
// adjacency: scipy cdist 'euclidean' on f32 inputs upcast to f64, |d_prev - d_new| <= thr.
// One block per problem; thread t owns (row i, word w) pairs.  Self-loops are dropped
// (networkx: adj[u] = {v for v in G[u] if v != u}).
// ------------------------------------------------------------------------------------
// T = float: the f32 coordinates cv2 returns, widened like scipy's cdist does; T = double: callers that hold float64
// coordinates (metric or undistorted points), compared in full precision as the reference would
template <typename T>
__global__ void __launch_bounds__(256)
k_adjacency(const T* __restrict__ prev, const T* __restrict__ nw, const int32_t* __restrict__ counts, int Kstride,
            int Kpad, int NW, double thr, uint32_t* __restrict__ adjbits, uint8_t* __restrict__ adj_bytes) {
    const int p = blockIdx.x;
    const int K = counts[p];
    const T* a = prev + (size_t)p * Kstride * 2;
    const T* b = nw + (size_t)p * Kstride * 2;
    uint32_t* out = adjbits + (size_t)p * Kpad * NW;
    for (int t = threadIdx.x; t < Kpad * NW; t += blockDim.x) {
        const int i = t / NW, w = t - i * NW;
        uint32_t bitsv = 0;
        if (i < K) {
            const double ax = (double)a[2 * i], ay = (double)a[2 * i + 1];
            const double bx = (double)b[2 * i], by = (double)b[2 * i + 1];
            for (int bb = 0; bb < 32; ++bb) {
                const int j = w * 32 + bb;
                if (j >= K) break;
                const double d1x = __dsub_rn(ax, (double)a[2 * j]), d1y = __dsub_rn(ay, (double)a[2 * j + 1]);
                const double d2x = __dsub_rn(bx, (double)b[2 * j]), d2y = __dsub_rn(by, (double)b[2 * j + 1]);
                const double s1 = __dadd_rn(__dmul_rn(d1x, d1x), __dmul_rn(d1y, d1y));
                const double s2 = __dadd_rn(__dmul_rn(d2x, d2x), __dmul_rn(d2y, d2y));
                const double dd = fabs(__dsub_rn(__dsqrt_rn(s1), __dsqrt_rn(s2)));
                const bool e = dd <= thr;
                if (adj_bytes) adj_bytes[((size_t)p * K + i) * K + j] = e;   // includes the diagonal, like the reference matrix
                if (e && j != i) bitsv |= 1u << bb;
            }
        }
        out[t] = bitsv;
    }
}

__global__ void __launch_bounds__(256)
k_bytes_to_bits(const uint8_t* __restrict__ adj, int K, int Kpad, int NW, uint32_t* __restrict__ out) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < Kpad * NW; t += gridDim.x * blockDim.x) {
        const int i = t / NW, w = t - i * NW;
        uint32_t v = 0;
        if (i < K)
            for (int bb = 0; bb < 32; ++bb) {
                const int j = w * 32 + bb;
                if (j < K && j != i && adj[(size_t)i * K + j]) v |= 1u << bb;
            }
        out[t] = v;
    }
}

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
struct CliqueWorkspace {
    CliqueGeom g;
    int P;
    uint32_t* adjbits; int16_t* adjseq; uint32_t* stack;
    uint8_t* mask; int32_t *n_inliers, *nodes, *status; long long* n_yields; unsigned long long* hash;
    int32_t* max_size; uint32_t* mc_levels; uint32_t* mc_arena; int32_t* mc_meta;
};

static CliqueGeom make_geom(int Kmax) {
    CliqueGeom g;
    int kp = 64; while (kp < Kmax) kp <<= 1;
    g.Kpad = kp; g.NW = kp / 32; g.TABN = kp / 2; g.SW = g.NW + 4 + g.TABN / 2;
    g.SEQCAP = kp <= 512 ? 76 : (kp <= 2048 ? 306 : 1228);
    if (g.SEQCAP > kp) g.SEQCAP = kp;
    g.RS = g.NW | 1;
    return g;
}

size_t rf_clique_workspace_bytes(int Kmax, int P) {
    CliqueGeom g = make_geom(Kmax);
    size_t per = (size_t)g.Kpad * g.NW * 4 + (size_t)g.Kpad * g.SEQCAP * 2 +
                 (size_t)(g.Kpad + 1) * (3 * g.SW + 1) * 4 + (size_t)g.Kpad + 64;
    return per * P + 4096;
}

// carve a workspace out of `base` (device memory of rf_clique_workspace_bytes)
static CliqueWorkspace carve(void* base, int Kmax, int P) {
    CliqueWorkspace ws; ws.g = make_geom(Kmax); ws.P = P;
    const CliqueGeom& g = ws.g;
    char* p = (char*)base;
    auto take = [&](size_t bytes) { char* r = p; p += (bytes + 255) & ~(size_t)255; return r; };
    ws.stack = (uint32_t*)take((size_t)P * (g.Kpad + 1) * (3 * g.SW + 1) * 4);
    ws.adjbits = (uint32_t*)take((size_t)P * g.Kpad * g.NW * 4);
    ws.adjseq = (int16_t*)take((size_t)P * g.Kpad * g.SEQCAP * 2);
    ws.mask = (uint8_t*)take((size_t)P * g.Kpad);
    ws.n_inliers = (int32_t*)take((size_t)P * 4);
    ws.nodes = (int32_t*)take((size_t)P * 4);
    ws.status = (int32_t*)take((size_t)P * 4);
    ws.n_yields = (long long*)take((size_t)P * 8);
    ws.hash = (unsigned long long*)take((size_t)P * 8);
    ws.max_size = (int32_t*)take((size_t)P * 4);
    ws.mc_levels = (uint32_t*)take((size_t)P * MC_WARPS * g.Kpad * g.NW * 4);
    ws.mc_arena = (uint32_t*)take((size_t)P * MC_WARPS * MC_ARENA * 4);
    ws.mc_meta = (int32_t*)take((size_t)P * MC_WARPS * g.Kpad * 2 * 4);
    return ws;
}

size_t rf_clique_ws_total(int Kmax, int P) {
    CliqueGeom g = make_geom(Kmax);
    auto r = [](size_t b) { return (b + 255) & ~(size_t)255; };
    return r((size_t)P * (g.Kpad + 1) * (3 * g.SW + 1) * 4) + r((size_t)P * g.Kpad * g.NW * 4) + r((size_t)P * g.Kpad * g.SEQCAP * 2) +
           r((size_t)P * g.Kpad) + 3 * r((size_t)P * 4) + 2 * r((size_t)P * 8) +
           r((size_t)P * 4) + r((size_t)P * MC_WARPS * g.Kpad * g.NW * 4) + r((size_t)P * MC_WARPS * MC_ARENA * 4) +
           r((size_t)P * MC_WARPS * g.Kpad * 2 * 4);
}

static int launch_clique(rf_handle* h, const CliqueWorkspace& ws, const int32_t* d_counts, int prune, bool debug) {
    CliqueArgs a;
    a.g = ws.g; a.P = ws.P; a.counts = d_counts; a.Kmax = ws.g.Kpad; a.adjbits = ws.adjbits; a.adjseq = ws.adjseq;
    a.stack = ws.stack; a.prune = prune; a.node_limit = h->cfg.clique_node_limit;
    a.mask = ws.mask; a.n_inliers = ws.n_inliers; a.nodes = ws.nodes; a.status = ws.status;
    a.n_yields = debug ? ws.n_yields : nullptr; a.order_hash = debug ? ws.hash : nullptr;
    a.prof = nullptr;
    a.max_size = nullptr;
    a.vstar = nullptr; a.vstar_stride = 0;
    static const bool no_mc = getenv("RF_CLIQUE_NO_MAXSIZE") != nullptr;   // diagnostic: the round-1 search (greedy bound only)
    if (!debug && prune && ws.g.Kpad <= 1024 && !no_mc) {
        MaxCliqueArgs m;
        m.P = ws.P; m.Kpad = ws.g.Kpad; m.NW = ws.g.NW; m.counts = d_counts; m.adjbits = ws.adjbits; m.max_size = ws.max_size;
        m.levels = ws.mc_levels; m.arena = ws.mc_arena; m.lvl_meta = ws.mc_meta;
        const int Kp = ws.g.Kpad, NWp = ws.g.NW;
        const size_t msm = (size_t)Kp * (NWp | 1) * 4 + (size_t)Kp * 2 * 2 + 8 + (size_t)Kp * 4;
        if (msm > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(k_maxclique, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msm);
            if (e != cudaSuccess) return rf_fail(h, RF_E_CUDA, "maxclique smem %zu: %s", msm, cudaGetErrorString(e));
        }
        k_maxclique<<<ws.P, 32 * MC_WARPS, msm, h->stream>>>(m);
        RF_CHECK_LAUNCH(h);
        a.max_size = ws.max_size;
        static const bool no_vi = getenv("RF_CLIQUE_NO_VIABLE") != nullptr;   // diagnostic: contextual bounds only
        const size_t vsm = (size_t)Kp * (NWp | 1) * 4;
        if (!no_vi && vsm <= 40 * 1024) {
            // V* lives in k_maxclique's level scratch, which is free once that kernel has finished
            const size_t vstride = (size_t)MC_WARPS * Kp * NWp;
            k_viable<<<ws.P, 32 * VI_WARPS, vsm, h->stream>>>(ws.g, ws.P, d_counts, ws.adjbits, ws.max_size, ws.mc_levels, vstride);
            RF_CHECK_LAUNCH(h);
            a.vstar = ws.mc_levels; a.vstar_stride = vstride;
        }
    }
    a.adjseq_ready = 0;
    static const bool no_as = getenv("RF_CLIQUE_NO_ADJSEQ_KERNEL") != nullptr;   // diagnostic: the in-search prologue
    {
        const int own_words = 2 * ws.g.NW > 128 ? 2 * ws.g.NW : 128;
        const size_t asm_b = (size_t)ADJSEQ_WARPS * (ws.g.Kpad + ws.g.TABN / 2 + own_words) * 4;
        if (!no_as && asm_b <= 48 * 1024) {
            k_adjseq<<<dim3(ADJSEQ_CTAS, ws.P), 32 * ADJSEQ_WARPS, asm_b, h->stream>>>(ws.g, ws.P, d_counts, ws.adjbits, ws.adjseq);
            RF_CHECK_LAUNCH(h);
            a.adjseq_ready = 1;
        }
    }
    static const bool want_prof = getenv("RF_CLIQUE_PROFILE") != nullptr;
    long long* d_prof = nullptr;
    if (want_prof && cudaMalloc(&d_prof, (size_t)ws.P * 128) == cudaSuccess) a.prof = d_prof;
    size_t smem = (size_t)5 * ws.g.SW * 4 + (size_t)5 * ws.g.Kpad * 2 + (size_t)(2 * ws.g.NW > 128 ? 2 * ws.g.NW : 128) * 4 + (size_t)ws.g.NW * 4;
    const size_t adj_bytes = (size_t)ws.g.Kpad * ws.g.RS * 4;
    a.adj_in_smem = smem + adj_bytes <= 96 * 1024;      // K <= 512: rows live next to the search frame
    if (a.adj_in_smem) smem += adj_bytes;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_clique, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return rf_fail(h, RF_E_CUDA, "clique smem %zu: %s", smem, cudaGetErrorString(e));
    }
    k_clique<<<ws.P, 32, smem, h->stream>>>(a);
    RF_CHECK_LAUNCH(h);
    if (d_prof) {   // diagnostic only: synchronous dump of the per-phase cycle counters
        std::vector<long long> hp((size_t)ws.P * 16);
        cudaMemcpyAsync(hp.data(), d_prof, hp.size() * 8, cudaMemcpyDeviceToHost, h->stream);
        cudaStreamSynchronize(h->stream);
        cudaFree(d_prof);
        static const char* names[8] = {"stage", "deg+adjseq", "greedy", "pivot", "ext", "pop", "children", "frames"};
        long long tot[8] = {0}, mx = 0; int arg = 0;
        for (int p = 0; p < ws.P; ++p) { long long t = 0; for (int k = 0; k < 8; ++k) { tot[k] += hp[p * 16 + k]; t += hp[p * 16 + k]; } if (t > mx) { mx = t; arg = p; } }
        fprintf(stderr, "[clique profile] P=%d  mean cycles:", ws.P);
        for (int k = 0; k < 8; ++k) fprintf(stderr, " %s=%lld", names[k], tot[k] / ws.P);
        int kk = 0, nn = 0, ni = 0, ms = 0;
        cudaMemcpy(&kk, d_counts + arg, 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(&nn, ws.nodes + arg, 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(&ni, ws.n_inliers + arg, 4, cudaMemcpyDeviceToHost);
        if (a.max_size) cudaMemcpy(&ms, ws.max_size + arg, 4, cudaMemcpyDeviceToHost);
        fprintf(stderr, "\n[clique profile] slowest pair %d (%lld cycles; K=%d pops=%d clique=%d maxsize=%d):", arg, mx, kk, nn, ni, ms);
        for (int k = 0; k < 8; ++k) fprintf(stderr, " %s=%lld", names[k], hp[arg * 16 + k]);
        static const char* enames[8] = {"entries", "chains", "chain_vertices", "clique_shortcuts", "pivots", "descents", "bound_prunes", "leaf_yields"};
        fprintf(stderr, "\n[clique profile] slowest pair events:");
        for (int k = 0; k < 8; ++k) fprintf(stderr, " %s=%lld", enames[k], hp[arg * 16 + 8 + k]);
        fprintf(stderr, "\n");
    }
    return RF_OK;
}

// Batched device entry used by the fused pair/batch path.  d_prev/d_new: [P][Kstride][2]
// compacted good correspondences, d_counts[P].  `ws_base` must hold rf_clique_ws_total().
int rf_launch_reject(rf_handle* h, void* ws_base, const float* d_prev, const float* d_new, const int32_t* d_counts,
                     int Kstride, int P, uint8_t** d_mask_out, int* mask_stride, int32_t** d_ninl, int32_t** d_nodes,
                     int32_t** d_status) {
    CliqueWorkspace ws = carve(ws_base, Kstride, P);
    k_adjacency<float><<<P, 256, 0, h->stream>>>(d_prev, d_new, d_counts, Kstride, ws.g.Kpad, ws.g.NW, h->cfg.dist_thr_px,
                                          ws.adjbits, nullptr);
    RF_CHECK_LAUNCH(h);
    int rc = launch_clique(h, ws, d_counts, 1, false);
    if (rc) return rc;
    *d_mask_out = ws.mask; *mask_stride = ws.g.Kpad; *d_ninl = ws.n_inliers; *d_nodes = ws.nodes; *d_status = ws.status;
    return RF_OK;
}

template <typename T>
static int consistency_adjacency_impl(rf_handle* h, const T* prev_xy, const T* new_xy, int K, uint8_t* adj) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !prev_xy || !new_xy || !adj || K < 0) return rf_fail(h, RF_E_BADARG, "rf_consistency_adjacency: bad argument");
    if (K == 0) return RF_OK;
    CliqueGeom g = make_geom(K);
    size_t bp = ((size_t)K * 2 * sizeof(T) + 255) & ~(size_t)255;
    size_t total = 2 * bp + 256 + (size_t)g.Kpad * g.NW * 4 + (size_t)K * K;
    int rc = rf_ensure_scratch(h, total);
    if (rc) return rc;
    char* base = (char*)h->d_scratch;
    T* dp = (T*)base; T* dn = (T*)(base + bp);
    int32_t* dc = (int32_t*)(base + 2 * bp);
    uint32_t* bits = (uint32_t*)(base + 2 * bp + 256);
    uint8_t* bytes = (uint8_t*)(bits + (size_t)g.Kpad * g.NW);
    RF_CUDA(h, cudaMemcpyAsync(dp, prev_xy, (size_t)K * 2 * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    RF_CUDA(h, cudaMemcpyAsync(dn, new_xy, (size_t)K * 2 * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    RF_CUDA(h, cudaMemcpyAsync(dc, &K, 4, cudaMemcpyHostToDevice, h->stream));
    k_adjacency<T><<<1, 256, 0, h->stream>>>(dp, dn, dc, K, g.Kpad, g.NW, h->cfg.dist_thr_px, bits, bytes);
    RF_CHECK_LAUNCH(h);
    RF_CUDA(h, cudaMemcpyAsync(adj, bytes, (size_t)K * K, cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaStreamSynchronize(h->stream));
    return RF_OK;
}

template <typename T>
static int reject_outliers_impl(rf_handle* h, const T* prev_xy, const T* new_xy, int K, uint8_t* mask, int* n_inliers,
                                int* nodes) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !prev_xy || !new_xy || !mask || K < 0) return rf_fail(h, RF_E_BADARG, "rf_reject_outliers: bad argument");
    if (n_inliers) *n_inliers = 0;
    if (nodes) *nodes = 0;
    if (K == 0) return RF_OK;
    if (K > 8192) return rf_fail(h, RF_E_CAPACITY, "rf_reject_outliers: K=%d exceeds 8192", K);
    size_t bp = ((size_t)K * 2 * sizeof(T) + 255) & ~(size_t)255;
    size_t wsb = rf_clique_ws_total(K, 1);
    int rc = rf_ensure_scratch(h, 2 * bp + 256 + wsb);
    if (rc) return rc;
    char* base = (char*)h->d_scratch;
    T* dp = (T*)base; T* dn = (T*)(base + bp);
    int32_t* dc = (int32_t*)(base + 2 * bp);
    RF_CUDA(h, cudaMemcpyAsync(dp, prev_xy, (size_t)K * 2 * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    RF_CUDA(h, cudaMemcpyAsync(dn, new_xy, (size_t)K * 2 * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    RF_CUDA(h, cudaMemcpyAsync(dc, &K, 4, cudaMemcpyHostToDevice, h->stream));
    uint8_t* dmask; int stride; int32_t *dn_inl, *dnodes, *dstatus;
    {
        CliqueWorkspace ws = carve(base + 2 * bp + 256, K, 1);
        k_adjacency<T><<<1, 256, 0, h->stream>>>(dp, dn, dc, K, ws.g.Kpad, ws.g.NW, h->cfg.dist_thr_px, ws.adjbits, nullptr);
        RF_CHECK_LAUNCH(h);
        rc = launch_clique(h, ws, dc, 1, false);
        if (rc) return rc;
        dmask = ws.mask; stride = ws.g.Kpad; dn_inl = ws.n_inliers; dnodes = ws.nodes; dstatus = ws.status;
    }
    int32_t out[3];
    RF_CUDA(h, cudaMemcpyAsync(mask, dmask, (size_t)K, cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaMemcpyAsync(&out[0], dn_inl, 4, cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaMemcpyAsync(&out[1], dnodes, 4, cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaMemcpyAsync(&out[2], dstatus, 4, cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaStreamSynchronize(h->stream));
    if (n_inliers) *n_inliers = out[0];
    if (nodes) *nodes = out[1];
    if (out[2] != RF_OK) return rf_fail(h, out[2], "rf_reject_outliers: clique search exceeded %lld nodes (best-so-far mask returned)",
                                        (long long)h->cfg.clique_node_limit);
    return RF_OK;
}

extern "C" {

int rf_consistency_adjacency(rf_handle* h, const float* prev_xy, const float* new_xy, int K, uint8_t* adj) {
    return consistency_adjacency_impl<float>(h, prev_xy, new_xy, K, adj);
}
int rf_consistency_adjacency_f64(rf_handle* h, const double* prev_xy, const double* new_xy, int K, uint8_t* adj) {
    return consistency_adjacency_impl<double>(h, prev_xy, new_xy, K, adj);
}
int rf_reject_outliers(rf_handle* h, const float* prev_xy, const float* new_xy, int K, uint8_t* mask, int* n_inliers, int* nodes) {
    return reject_outliers_impl<float>(h, prev_xy, new_xy, K, mask, n_inliers, nodes);
}
int rf_reject_outliers_f64(rf_handle* h, const double* prev_xy, const double* new_xy, int K, uint8_t* mask, int* n_inliers, int* nodes) {
    return reject_outliers_impl<double>(h, prev_xy, new_xy, K, mask, n_inliers, nodes);
}

// Test hook: clique search on a caller-supplied adjacency matrix (K x K bytes).
int rf_clique_search(rf_handle* h, const uint8_t* adj, int K, int prune, int32_t* clique_mask_out, int* size,
                     int64_t* n_yields, uint64_t* order_hash, int64_t* nodes) {
    RfDeviceGuard rf_guard_(h);
    if (!h || !adj || K < 0) return rf_fail(h, RF_E_BADARG, "rf_clique_search: bad argument");
    if (K == 0) { if (size) *size = 0; return RF_OK; }
    size_t ab = ((size_t)K * K + 255) & ~(size_t)255;
    size_t wsb = rf_clique_ws_total(K, 1);
    int rc = rf_ensure_scratch(h, ab + 256 + wsb);
    if (rc) return rc;
    char* base = (char*)h->d_scratch;
    uint8_t* dadj = (uint8_t*)base;
    int32_t* dc = (int32_t*)(base + ab);
    CliqueWorkspace ws = carve(base + ab + 256, K, 1);
    RF_CUDA(h, cudaMemcpyAsync(dadj, adj, (size_t)K * K, cudaMemcpyHostToDevice, h->stream));
    RF_CUDA(h, cudaMemcpyAsync(dc, &K, 4, cudaMemcpyHostToDevice, h->stream));
    k_bytes_to_bits<<<64, 256, 0, h->stream>>>(dadj, K, ws.g.Kpad, ws.g.NW, ws.adjbits);
    RF_CHECK_LAUNCH(h);
    rc = launch_clique(h, ws, dc, prune & 1, !(prune & 2));
    if (rc) return rc;
    std::vector<uint8_t> m(K);
    int32_t o[3]; long long ny; unsigned long long hs;
    RF_CUDA(h, cudaMemcpyAsync(m.data(), ws.mask, (size_t)K, cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaMemcpyAsync(&o[0], ws.n_inliers, 4, cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaMemcpyAsync(&o[1], ws.nodes, 4, cudaMemcpyDeviceToHost, h->stream));
    RF_CUDA(h, cudaMemcpyAsync(&o[2], ws.status, 4, cudaMemcpyDeviceToHost, h->stream));
    ny = 0; hs = 0;
    if (!(prune & 2)) {
        RF_CUDA(h, cudaMemcpyAsync(&ny, ws.n_yields, 8, cudaMemcpyDeviceToHost, h->stream));
        RF_CUDA(h, cudaMemcpyAsync(&hs, ws.hash, 8, cudaMemcpyDeviceToHost, h->stream));
    }
    RF_CUDA(h, cudaStreamSynchronize(h->stream));
    if (clique_mask_out) for (int i = 0; i < K; ++i) clique_mask_out[i] = m[i];
    if (size) *size = o[0];
    if (nodes) *nodes = o[1];
    if (n_yields) *n_yields = ny;
    if (order_hash) *order_hash = hs;
    if (o[2] != RF_OK) return rf_fail(h, o[2], "rf_clique_search: node limit exceeded");
    return RF_OK;
}

}  // extern "C"
