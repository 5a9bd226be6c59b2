// rg_plan.hpp -- host-side planner of the convolution schedule (pure C++, no CUDA).
//
// The Poisson bracket of RG.liouville (RG.py:115-117) is two truncated linear convolutions over
// (j, nu) (RG.conv_product, RG.py:102-106).  On the device one CTA owns one Hamiltonian; each warp
// owns up to 32 output modes nu (one per lane, all J+1 Taylor accumulators in registers) and walks a
// warp-uniform list of source modes nu1 of the generator y (broadcast from shared memory) while each
// lane gathers g[nu - nu1].  This file decides
//   * which output modes share a warp (compact boxes keep the in-range sets R(nu) similar),
//   * each warp's list of nu1 (restricted to the I^- support of y, RG.py:77-86,108-113),
//   * a static split of the heaviest lists onto the lightest warps (<= 2 pieces per box so the
//     partial sums combine in a fixed order -- deterministic, no atomics).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <numeric>
#include <vector>

namespace rgplan {

struct Geometry {
    int dim = 0, L = 0, J = 0, JP = 0, n1 = 0, M = 0, vecjl = 0;
    int off2 = 0;       // record layout: [0..J-1] d-type (dA g or dA y), [off2..off2+J] Omega-type
    int rec_units = 0;  // 16-byte units per record (odd => conflict-free LDS.128 over consecutive modes)
    int g_units = 0;    // 16-byte units of a g-only record (g[0..J], product-rule form of the BIG kernels); odd too
    int stride[4] = {0, 0, 0, 0};
    void init(int dim_, int L_, int J_) {
        dim = dim_; L = L_; J = J_; JP = J + 1; n1 = 2 * L + 1;
        M = 1; for (int d = 0; d < dim; ++d) M *= n1;
        vecjl = JP * M;
        off2 = J + (J & 1);
        int tot = off2 + JP;
        rec_units = (tot + 1) / 2;
        if ((rec_units & 1) == 0) rec_units += 1;
        g_units = (JP + 1) / 2;
        if ((g_units & 1) == 0) g_units += 1;
        int s = 1;
        for (int d = dim - 1; d >= 0; --d) { stride[d] = s; s *= n1; }
        for (int d = dim; d < 4; ++d) stride[d] = 0;
    }
    // centred coords (biased by +L, each in [0,2L]) <-> centred linear index
    void coords(int c, int* v) const { for (int d = 0; d < dim; ++d) { v[d] = (c / stride[d]) % n1; } for (int d = dim; d < 4; ++d) v[d] = 0; }
    int wrapped_of_centred(int c) const {
        int v[4]; coords(c, v); int w = 0;
        for (int d = 0; d < dim; ++d) { int nu = v[d] - L; int wi = nu >= 0 ? nu : nu + n1; w = w * n1 + wi; }
        return w;
    }
    int centred_of_wrapped(int w) const {
        int c = 0, tmp[4];
        for (int d = dim - 1; d >= 0; --d) { tmp[d] = w % n1; w /= n1; }
        for (int d = 0; d < dim; ++d) { int nu = tmp[d] <= L ? tmp[d] : tmp[d] - n1; c = c * n1 + (nu + L); }
        return c;
    }
    uint32_t pack(const int* v) const { return (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16); }
};

// FMAs of one (nu1, nu2) pair for a source mode with jm non-zero Taylor levels (RG.py:115-117):
//   conv(dA y, Omega g): j1 <= J-1, needs y[j1+1] != 0  -> j1+1 < jm
//   conv(Omega y, dA g): j1 <= J,   needs y[j1]   != 0  -> j1   < jm ; dA g has J levels
inline int taylor_fma(int jm, int J) {
    int c = 0;
    for (int j1 = 0; j1 < J; ++j1) if (j1 + 1 < jm) c += J - j1 + 1;
    for (int j1 = 0; j1 <= J; ++j1) if (j1 < jm) c += std::min(J, J - j1 + 1);
    return c;
}

// FMAs the device ISSUES for such a pair in the product-rule form of the bracket (BIG kernels; rg_device.cuh, taylor_leib):
//   S1: y1[i] g[b], i + 1 < jm, b <= J - i;   S2: y2[a] g[b], a < jm, a + b - 1 in [0, J]      (140 instead of 130 at J = 10)
inline int taylor_fma_issued_pr(int jm, int J) {
    int c = 0;
    for (int i = 0; i < J; ++i) if (i + 1 < jm) c += J - i + 1;
    for (int a = 0; a <= J; ++a) if (a < jm) c += (a == 0) ? J : J + 2 - a;
    return c;
}

// Every list entry runs a branch-free pipelined Taylor product on its zero-padded y record, truncated at compile
// time to LIM_A, LIM_B or all J+1 levels of y (measured: the early-exit path costs a complete product whatever
// the number of levels).  Returns the number of levels the device multiplies for an entry with jm levels present
// (jm itself on the early-exit path).  Mirrors RecGeom<J>::LIM_A / LIM_B and the dispatch in conv_piece.
inline int level_class(int jm, int J) {
    if (jm > J + 1) return jm;          // caller-supplied generator: early-exit path
    const int lim_a = (J >= 8) ? 3 : (J / 2 > 2 ? J / 2 : 2), lim_b = (J >= 8) ? 6 : J - 1;
    return jm > lim_b ? J + 1 : (jm > lim_a ? lim_b : lim_a);
}

struct Piece { int box, begin, end, rank; };  // rank: position of the piece within its box's list (0 = head)

constexpr int MAXP = 4;    // least number of piece slots per warp and stage: any number of heads (rank 0), at most one continuation
constexpr int MAXSTAGE = 2;

// A schedule has one or two STAGES.  A stage is a partition of the output modes into boxes of 32 lanes plus,
// for every box, the list of source modes it visits.  With two stages every source mode of y is visited in the
// stage whose box shape wastes fewer lanes on it (in-range sets of a truncated convolution differ from lane to
// lane; tall boxes suit sources whose range test bites along the last axis, wide boxes the others); stage 0
// stores its sums, stage 1 adds on top behind a barrier.
struct Schedule {
    int n_warps = 0;    // warps of the cluster (csize CTAs x warps per CTA); pieces are dealt to these
    int csize = 1;      // CTAs per Hamiltonian
    int n_stages = 1;
    int maxp = MAXP;    // piece slots per warp and stage (grows with the number of boxes a warp has to own)
    int fin_begin[4] = {0, 0, 0, 0}, fin_end[4] = {0, 0, 0, 0};   // boxes owned (computed and broadcast) by each CTA (1 stage only)
    int stage0_pieces[4] = {0, 0, 0, 0};                          // pieces each CTA combines in the storing stage
    int n_boxes = 0;    // boxes of all stages
    int n_support = 0;
    int max_rank[MAXSTAGE] = {0, 0};
    bool half = false;  // outputs restricted to the half space nu <= 0 (exactly symmetric operands)
    std::vector<int32_t> lane_out;     // [NB*32] centred mode index or -1
    std::vector<uint32_t> lane_coord;  // [NB*32] packed biased coords (idle lanes: first coordinate 255)
    std::vector<Piece> pieces;         // [n_stages][NW*maxp], box = -1 for unused slots; heads first, continuation last
    std::vector<int32_t> piece_runs;   // [n_stages][NW*maxp][2]: end of the piece's run of complete products, end of its LIM_B run
    bool general = false;              // caller-supplied generator (dense): every entry takes the general product
    bool prule = false;                // costs counted for the product-rule form of the bracket (BIG kernels)
    std::vector<uint32_t> entries;     // [n*2]: {lanes of the piece's box whose nu - nu1 is on the grid (bit l = lane l), centred index of nu1 | rec<<16}
    std::vector<uint8_t> entry_jm;     // [n]: Taylor levels of y present at the entry's nu1 (J+2: caller-supplied generator)
    std::vector<int32_t> supp_rec;     // [M] record index of centred nu1, -1 if not in support
    std::vector<uint8_t> jm;           // [M] 1 + highest I^- level (0 = not in support)
    int64_t fma_useful = 0;            // per liouville, in-range & structurally non-zero
    int64_t fma_executed = 0;          // per liouville, 32 lanes x list cost
    int64_t cost_max = 0, cost_total = 0;  // per-warp list cost summed over stages (balance = total / (n_warps * max))
};

// mask[c] = bit j set iff (j, nu_c) in I^- (centred order).  dense: every mode, every level (jm = J+2 so
// that the y1[J] term of a caller-supplied generator is kept).  half: only modes c <= centre are computed.
// n_warps: warps of the whole cluster (csize CTAs x warps per CTA).
inline Schedule make_schedule_impl(const Geometry& g, const std::vector<uint16_t>& mask, bool dense, bool half, int n_warps, int csize,
                                   bool allow_two_stage, bool prule) {
    Schedule s;
    const int M = g.M, L = g.L, J = g.J, c0 = (M - 1) / 2;
    s.half = half;
    s.csize = csize;
    s.general = dense;
    s.prule = prule;
    s.jm.assign(M, 0);
    s.supp_rec.assign(M, -1);
    for (int c = 0; c < M; ++c) {
        int jm = 0;
        if (dense) jm = g.JP + 1;
        else { uint16_t m = mask[c]; while (m) { ++jm; m >>= 1; } }
        s.jm[c] = (uint8_t)jm;
        if (jm > 0) s.supp_rec[c] = s.n_support++;
    }
    auto tcost = [&](int c1) { return (int64_t)taylor_fma(std::min<int>(s.jm[c1], g.JP), J) + (s.jm[c1] > g.JP ? 1 : 0); };
    auto tissued = [&](int c1) {
        if (s.jm[c1] > g.JP) return tcost(c1);
        return (int64_t)(prule ? taylor_fma_issued_pr(level_class(s.jm[c1], J), J) : taylor_fma(level_class(s.jm[c1], J), J));
    };
    // issue-slot model of one list entry, used to balance the warps: a DFMA holds the issue port ~2.65 cycles on
    // average (operand-reuse misses included), every entry carries ~67 other instructions (loads, decode, loop)
    auto wcost_of = [&](int c1) { return (int64_t)(2.65 * (double)tissued(c1)) + 67; };
    std::vector<int> outputs;
    for (int c = 0; c < M; ++c) if (!half || c <= c0) outputs.push_back(c);
    const int NO = (int)outputs.size();
    const int NBS = (NO + 31) / 32;   // boxes per stage
    const int last = g.dim - 1;
    // ---- candidate partitions (orders of the output modes, cut into boxes of 32 consecutive lanes).  In each a
    // quarter-warp gathers 8 records whose indices are distinct mod 8 (consecutive along the last axis, or a
    // stride of n1 = odd along the previous one) -> conflict-free LDS.128 with the odd record stride.
    //   0: (column group of 8 along the last axis, leading axes, column)  -- compact 4 x 8 boxes
    //   1: column-major (last axis slowest)                               -- tall boxes   [2-D only]
    //   2: (column group of 16, row, column)                              -- wide boxes   [2-D only]
    //   3: centred index order (lines along the last axis)                -- 1 x 1 x 32   [3-D]
    //   4: blocks of 4 x 1 x 8 (first, second, last axis)                 -- slabs        [3-D]
    // (3-D, L = J = 6 spiral mean: one partition of 1 x 4 x 8 boxes keeps 0.61 of the issued lanes in range, the
    //  pair 3 + 4 with every source mode in its cheaper partition 0.72; flat 8 x 4 x 1 blocks would give 0.76 but
    //  their gathers touch 32 separate cache lines per load when the records live in global memory.)
    auto make_order = [&](int kind) {
        std::vector<int> order = outputs;
        auto key = [&](int c) {
            int v[4]; g.coords(c, v);
            uint64_t k;
            if (kind == 3) return (uint64_t)c;
            if (kind == 4) return ((((uint64_t)(v[0] / 4) * 128 + (uint64_t)v[1]) * 128 + (uint64_t)(v[2] / 8)) * 4 + (uint64_t)(v[0] % 4)) * 8 + (uint64_t)(v[2] % 8);
            if (kind == 1) { k = (uint64_t)v[last]; for (int d = 0; d < last; ++d) k = k * 128 + (uint64_t)v[d]; return k; }
            k = (uint64_t)(v[last] / (kind == 2 ? 16 : 8));
            for (int d = 0; d < last; ++d) k = k * 128 + (uint64_t)v[d];
            return k * 128 + (uint64_t)v[last];
        };
        std::sort(order.begin(), order.end(), [&](int a, int b) { return key(a) < key(b); });
        // Lane order inside a box.  A quarter-warp's LDS.128 over the gathered records G[nu - nu1] is conflict free
        // iff its eight record indices are distinct mod 8 (odd record stride in 16-byte units); nu - nu1 shifts all
        // indices of a box alike, so this is a property of the box.  Sort the box by (index mod 8) and deal the
        // sorted positions round-robin to the four quarters: a residue class of <= 4 modes never meets itself.
        // First make every full box hold at most four modes of each residue class: walk the sorted order and let
        // a mode whose class is already full in the current box wait for the next one (look-ahead of 24 modes, so
        // boxes keep their shape up to a few swapped cells).
        {
            std::vector<int> rest(order), res;
            res.reserve(order.size());
            while (!rest.empty()) {
                int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, taken = 0;
                std::vector<char> used(rest.size(), 0);
                const size_t want = std::min<size_t>(32, rest.size());
                for (size_t i = 0; i < rest.size() && i < 32 + 24 && (size_t)taken < want; ++i)
                    if (rest.size() < 32 || cnt[rest[i] & 7] < 4) { used[i] = 1; ++cnt[rest[i] & 7]; ++taken; }
                for (size_t i = 0; i < rest.size() && (size_t)taken < want; ++i)
                    if (!used[i]) { used[i] = 1; ++taken; }
                std::vector<int> keep;
                for (size_t i = 0; i < rest.size(); ++i) { if (used[i]) res.push_back(rest[i]); else keep.push_back(rest[i]); }
                rest.swap(keep);
            }
            order.swap(res);
        }
        for (size_t b0 = 0; b0 + 32 <= order.size(); b0 += 32) {
            std::vector<int> box(order.begin() + b0, order.begin() + b0 + 32);
            std::stable_sort(box.begin(), box.end(), [](int a, int b) { return (a & 7) < (b & 7); });
            for (int p = 0; p < 32; ++p) order[b0 + (p % 4) * 8 + p / 4] = box[p];
        }
        // The last, partial box (n < 32 modes on lanes 0..n-1: quarters of 8, ..., n mod 8 lanes): the members of a
        // residue class go to distinct quarters, the roomiest quarter first; only a class with more members than
        // quarters with room left meets itself.  (Without this the 29-mode tail box of the half-space schedule at
        // L = 10 paid a second wavefront on three of its four quarter-gathers: 9 % of all gather wavefronts.)
        {
            const size_t b0 = order.size() / 32 * 32, n = order.size() - b0;
            if (n > 0) {
                std::vector<int> cls[8], quarter[4];
                for (size_t i = 0; i < n; ++i) cls[order[b0 + i] & 7].push_back(order[b0 + i]);
                int room[4];
                for (int q = 0; q < 4; ++q) room[q] = (int)std::min<size_t>(8, n > (size_t)8 * q ? n - 8 * q : 0);
                int idx[8] = {0, 1, 2, 3, 4, 5, 6, 7};
                std::stable_sort(idx, idx + 8, [&](int a, int b) { return cls[a].size() > cls[b].size(); });
                std::vector<int> spill;
                for (int k = 0; k < 8; ++k) {
                    bool used[4] = {false, false, false, false};
                    for (int c : cls[idx[k]]) {
                        int best = -1;
                        for (int q = 0; q < 4; ++q) if (!used[q] && room[q] > 0 && (best < 0 || room[q] > room[best])) best = q;
                        if (best < 0) { spill.push_back(c); continue; }
                        used[best] = true; --room[best]; quarter[best].push_back(c);
                    }
                }
                for (int c : spill) for (int q = 0; q < 4; ++q) if (room[q] > 0) { --room[q]; quarter[q].push_back(c); break; }
                size_t p = b0;
                for (int q = 0; q < 4; ++q) for (int c : quarter[q]) order[p++] = c;
            }
        }
        return order;
    };
    // lanes of box b (bit l = lane l) whose output mode nu has nu - nu1 on the grid
    auto valid_mask = [&](const std::vector<int>& order, int b, int c1) {
        int v1[4]; g.coords(c1, v1);
        uint32_t m = 0;
        for (int l = 0; l < 32 && b * 32 + l < NO; ++l) {
            int v[4]; g.coords(order[b * 32 + l], v);
            bool ok = true;
            for (int d = 0; d < g.dim; ++d) { int u = v[d] - v1[d] + L; if (u < 0 || u > 2 * L) ok = false; }
            if (ok) m |= 1u << l;
        }
        return m;
    };
    auto valid_lanes = [&](const std::vector<int>& order, int b, int c1) { return __builtin_popcount(valid_mask(order, b, c1)); };
    std::vector<std::vector<int>> orders;
    std::vector<int> stage_of(M, 0);
    const bool two_stage = (allow_two_stage && (g.dim == 2 || g.dim == 3) && csize == 1 && NBS >= 2);
    if (two_stage) {
        orders.push_back(make_order(g.dim == 2 ? 1 : 3));
        orders.push_back(make_order(g.dim == 2 ? 2 : 4));
        int64_t n_second = 0;
        for (int c1 = 0; c1 < M; ++c1) {
            if (s.jm[c1] == 0) continue;
            int nb[2] = {0, 0};
            for (int k = 0; k < 2; ++k) for (int b = 0; b < NBS; ++b) nb[k] += valid_lanes(orders[k], b, c1) > 0;
            stage_of[c1] = nb[1] < nb[0] ? 1 : 0;
            n_second += stage_of[c1];
        }
        if (n_second == 0) orders.pop_back();
    } else {
        orders.push_back(make_order(0));
    }
    s.n_stages = (int)orders.size();
    s.n_boxes = NBS * s.n_stages;
    s.n_warps = n_warps > 0 ? n_warps : NBS;
    const int NW = s.n_warps, NWc = NW / csize;
    // piece slots per warp and stage: boxes are dealt by cost, and in three dimensions many boxes are cheap, so a
    // warp may own any number of them there
    s.maxp = (g.dim >= 3) ? NBS + 1 : std::max(MAXP, (NBS + NWc - 1) / NWc + 2);
    const int MP = s.maxp;
    s.lane_out.assign(s.n_boxes * 32, -1);
    s.lane_coord.assign(s.n_boxes * 32, 0xFFFFFFFFu);
    s.pieces.assign((size_t)s.n_stages * NW * MP, Piece{-1, 0, 0, 0});
    std::vector<int64_t> wcost(NW, 0);
    for (int st = 0; st < s.n_stages; ++st) {
        const std::vector<int>& order = orders[st];
        const int box0 = st * NBS;
        for (int i = 0; i < NO; ++i) {
            s.lane_out[box0 * 32 + i] = order[i];
            int v[4]; g.coords(order[i], v);
            s.lane_coord[box0 * 32 + i] = g.pack(v);
        }
        // Idle lanes of the partial last box: coordinates far outside the grid along the first axis (the lane always
        // gathers a zero record) and, along the last axis, such that the zero record it picks (index congruent mod 8
        // to the record it "would have read") is in a bank group none of the quarter's active lanes uses.
        for (int i = NO; i < NBS * 32; ++i) {
            const int q0 = i / 8 * 8;
            bool used[8] = {false, false, false, false, false, false, false, false};
            for (int l = q0; l < q0 + 8 && l < NO; ++l) used[order[l] & 7] = true;
            int r = 0;
            while (r < 7 && used[r]) ++r;
            int v[4] = {255, 0, 0, 0};
            v[g.dim - 1] = ((r - 255 * g.stride[0]) % 8 + 8) % 8;
            if (g.dim == 1) v[0] = 255;   // one axis only: no room for a residue, keep the old convention
            s.lane_coord[box0 * 32 + i] = g.pack(v);
        }
        // ---- per-box union lists of this stage's source modes
        std::vector<std::vector<int>> lists(NBS);
        std::vector<int64_t> bcost(NBS, 0);
        int64_t total = 0;
        for (int b = 0; b < NBS; ++b) {
            for (int c1 = 0; c1 < M; ++c1) {
                if (s.jm[c1] == 0 || stage_of[c1] != st) continue;
                const int nvalid = valid_lanes(order, b, c1);
                if (!nvalid) continue;
                lists[b].push_back(c1);
                const int64_t t = tcost(c1);
                bcost[b] += wcost_of(c1); total += wcost_of(c1);
                s.fma_useful += (int64_t)nvalid * t;
                s.fma_executed += 32LL * tissued(c1);
            }
            // entries with every Taylor level present first: the branch-free path then runs in long stretches
            std::stable_sort(lists[b].begin(), lists[b].end(), [&](int x, int y) { return level_class(s.jm[x], J) > level_class(s.jm[y], J); });
        }
        s.cost_total += total;
        // ---- cluster mode: whole boxes are first partitioned over the `csize` CTAs (contiguous groups, minimal
        // maximum cost) so that every partial sum of a box stays inside one CTA; only the final rows of a box
        // travel through distributed shared memory.
        std::vector<int> gbeg(csize + 1, NBS);
        gbeg[0] = 0;
        if (csize > 1) {
            int64_t lo = 0, hi = 0;
            for (int b = 0; b < NBS; ++b) { lo = std::max(lo, bcost[b]); hi += bcost[b]; }
            auto fits = [&](int64_t cap, std::vector<int>* cuts) {
                int parts = 1; int64_t cur = 0;
                if (cuts) { cuts->assign(csize + 1, NBS); (*cuts)[0] = 0; }
                for (int b = 0; b < NBS; ++b) {
                    if (cur + bcost[b] > cap && cur > 0) { if (parts >= csize) return false; if (cuts) (*cuts)[parts] = b; ++parts; cur = 0; }
                    cur += bcost[b];
                }
                return true;
            };
            while (lo < hi) { int64_t mid = (lo + hi) / 2; if (fits(mid, nullptr)) hi = mid; else lo = mid + 1; }
            fits(lo, &gbeg);
            gbeg[csize] = NBS;
        }
        for (int r = 0; r < 4; ++r) { s.fin_begin[r] = r < csize ? box0 + gbeg[r] : box0 + NBS; s.fin_end[r] = r < csize ? box0 + gbeg[r + 1] : box0 + NBS; }
        // ---- inside every CTA: deal the concatenated lists of its boxes to its warps in equal-cost segments.  A
        // warp's segment starts with at most one continuation of a box (rank > 0) followed by heads of later boxes.
        Piece* pieces = s.pieces.data() + (size_t)st * NW * MP;
        std::vector<int> npieces(NW, 0);
        for (int r = 0; r < csize; ++r) {
            int64_t gtotal = 0;
            for (int b = gbeg[r]; b < gbeg[r + 1]; ++b) gtotal += bcost[b];
            const int w0 = r * NWc, w1 = w0 + NWc;
            int w = w0;
            int64_t acc = 0;
            for (int b = gbeg[r]; b < gbeg[r + 1]; ++b) {
                if (st > 0 && lists[b].empty()) continue;   // an adding stage has nothing to do for this box
                int rank = 0;
                size_t k = 0;
                do {
                    // open a piece of box b on warp w
                    while (w < w1 - 1 && (npieces[w] >= MP || (rank > 0 && npieces[w] > 0) ||
                                          acc * NWc >= gtotal * (int64_t)(w - w0 + 1))) ++w;
                    Piece pc{box0 + b, (int)s.entries.size() / 2, 0, rank};
                    while (k < lists[b].size()) {
                        const int c1 = lists[b][k];
                        if (w < w1 - 1 && acc * NWc >= gtotal * (int64_t)(w - w0 + 1) && (int)s.entries.size() / 2 > pc.begin) break;
                        // word0: the lanes of the box in range for nu1 (the range test is the planner's, not the kernel's; the
                        // level class of an entry is in the piece's run ends); word1: centred linear index | record<<16
                        s.entries.push_back(valid_mask(order, b, c1));
                        s.entries.push_back((uint32_t)c1 | ((uint32_t)s.supp_rec[c1] << 16));
                        s.entry_jm.push_back(s.jm[c1]);
                        acc += wcost_of(c1); wcost[w] += wcost_of(c1);
                        ++k;
                    }
                    pc.end = (int)s.entries.size() / 2;
                    if (npieces[w] >= MP) { s.n_warps = -1; return s; }   // too many boxes for this CTA; reported by the caller
                    // heads first, the continuation (if any) last
                    if (rank == 0) {
                        int q = npieces[w];
                        if (q > 0 && pieces[w * MP + q - 1].rank > 0) { pieces[w * MP + q] = pieces[w * MP + q - 1]; pieces[w * MP + q - 1] = pc; }
                        else pieces[w * MP + q] = pc;
                    } else {
                        pieces[w * MP + npieces[w]] = pc;
                    }
                    ++npieces[w];
                    if (st == 0) ++s.stage0_pieces[r];
                    s.max_rank[st] = std::max(s.max_rank[st], rank);
                    ++rank;
                } while (k < lists[b].size());
            }
        }
    }
    for (int i = 0; i < NW; ++i) { s.cost_max = std::max(s.cost_max, wcost[i]); }
    // run ends of every piece: its list is sorted by level class (complete products, LIM_B, LIM_A)
    s.piece_runs.assign(s.pieces.size() * 2, 0);
    for (size_t i = 0; i < s.pieces.size(); ++i) {
        const Piece& pc = s.pieces[i];
        if (pc.box < 0) continue;
        int e_full = pc.begin, e_mid;
        auto cls = [&](int e) { return level_class((int)s.entry_jm[(size_t)e], J); };
        const int lim_a = (J >= 8) ? 3 : (J / 2 > 2 ? J / 2 : 2), lim_b = (J >= 8) ? 6 : J - 1;   // as in level_class
        while (e_full < pc.end && cls(e_full) > std::max(lim_a, lim_b)) ++e_full;    // complete products (and the general ones)
        e_mid = e_full;
        if (lim_b > lim_a) while (e_mid < pc.end && cls(e_mid) == lim_b) ++e_mid;
        for (int e = e_mid; e < pc.end; ++e) if (cls(e) > lim_a) { s.n_warps = -1; return s; }    // list not sorted: cannot happen
        s.piece_runs[2 * i] = e_full; s.piece_runs[2 * i + 1] = e_mid;
    }
    return s;
}

// In three dimensions the two-partition schedule is not always the better one (the 8 x 4 x 1 blocks fit a grid of
// 13^3 modes, not one of 9^3): build both and keep the one that issues fewer FMAs.
inline Schedule make_schedule(const Geometry& g, const std::vector<uint16_t>& mask, bool dense, bool half, int n_warps, int csize = 1,
                              bool prule = false) {
    Schedule two = make_schedule_impl(g, mask, dense, half, n_warps, csize, true, prule);
    if (g.dim < 3 || two.n_stages < 2) return two;
    Schedule one = make_schedule_impl(g, mask, dense, half, n_warps, csize, false, prule);
    if (two.n_warps < 0) return one;
    if (one.n_warps < 0) return two;
    return one.fma_executed <= two.fma_executed ? one : two;
}

}  // namespace rgplan
