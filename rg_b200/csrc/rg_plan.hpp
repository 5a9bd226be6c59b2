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
    int stride[4] = {0, 0, 0, 0};
    void init(int dim_, int L_, int J_) {
        dim = dim_; L = L_; J = J_; JP = J + 1; n1 = 2 * L + 1;
        M = 1; for (int d = 0; d < dim; ++d) M *= n1;
        vecjl = JP * M;
        off2 = J + (J & 1);
        int tot = off2 + JP;
        rec_units = (tot + 1) / 2;
        if ((rec_units & 1) == 0) rec_units += 1;
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

struct Piece { int box_warp, begin, end, mode; };  // mode: 0 none, 1 store, 2 add (after barrier)

struct Schedule {
    int n_warps = 0;
    int n_support = 0;
    std::vector<int32_t> lane_out;     // [NW*32] centred mode index or -1
    std::vector<uint32_t> lane_coord;  // [NW*32] packed biased coords (0xFFFFFFFF idle)
    std::vector<Piece> pieces;         // [NW*2]
    std::vector<uint32_t> entries;     // [n*2]: {packed biased coords of nu1, (rec<<8)|jm}
    std::vector<int32_t> supp_rec;     // [M] record index of centred nu1, -1 if not in support
    std::vector<uint8_t> jm;           // [M] 1 + highest I^- level (0 = not in support)
    int64_t fma_useful = 0;            // per liouville, in-range & structurally non-zero
    int64_t fma_executed = 0;          // per liouville, 32 lanes x list cost
};

// mask[c] = bit j set iff (j, nu_c) in I^- (centred order).  dense: every mode, every level.
inline Schedule make_schedule(const Geometry& g, const std::vector<uint16_t>& mask, bool dense) {
    Schedule s;
    const int M = g.M, L = g.L, J = g.J;
    s.jm.assign(M, 0);
    s.supp_rec.assign(M, -1);
    for (int c = 0; c < M; ++c) {
        int jm = 0;
        if (dense) jm = g.JP;
        else { uint16_t m = mask[c]; while (m) { ++jm; m >>= 1; } }
        s.jm[c] = (uint8_t)jm;
        if (jm > 0) s.supp_rec[c] = s.n_support++;
    }
    // ---- output order: compact boxes, chunked by 32
    int box[4] = {1, 1, 1, 1};
    if (g.dim == 1) { box[0] = 32; }
    else if (g.dim == 2) { box[0] = 4; box[1] = 8; }
    else if (g.dim == 3) { box[0] = 2; box[1] = 4; box[2] = 4; }
    else { box[g.dim - 1] = 8; box[g.dim - 2] = 4; }
    std::vector<int> order(M);
    std::iota(order.begin(), order.end(), 0);
    auto key = [&](int c) {
        int v[4]; g.coords(c, v);
        uint64_t k = 0;
        for (int d = 0; d < g.dim; ++d) k = k * 64 + (uint64_t)(v[d] / box[d]);
        for (int d = 0; d < g.dim; ++d) k = k * 64 + (uint64_t)(v[d] % box[d]);
        return k;
    };
    std::sort(order.begin(), order.end(), [&](int a, int b) { return key(a) < key(b); });
    const int NW = (M + 31) / 32;
    s.n_warps = NW;
    s.lane_out.assign(NW * 32, -1);
    s.lane_coord.assign(NW * 32, 0xFFFFFFFFu);
    for (int i = 0; i < M; ++i) {
        s.lane_out[i] = order[i];
        int v[4]; g.coords(order[i], v);
        s.lane_coord[i] = g.pack(v);
    }
    // ---- per-warp union lists
    std::vector<std::vector<int>> lists(NW);
    std::vector<int64_t> cost(NW, 0);
    for (int w = 0; w < NW; ++w) {
        for (int c1 = 0; c1 < M; ++c1) {
            if (s.jm[c1] == 0) continue;
            int v1[4]; g.coords(c1, v1);
            int nvalid = 0;
            for (int l = 0; l < 32; ++l) {
                int co = s.lane_out[w * 32 + l];
                if (co < 0) continue;
                int v[4]; g.coords(co, v);
                bool ok = true;
                for (int d = 0; d < g.dim; ++d) { int u = v[d] - v1[d] + L; if (u < 0 || u > 2 * L) ok = false; }
                nvalid += ok;
            }
            if (nvalid) {
                lists[w].push_back(c1);
                int t = taylor_fma(s.jm[c1], J);
                cost[w] += t;
                s.fma_useful += (int64_t)nvalid * t;
                s.fma_executed += 32LL * t;
            }
        }
    }
    // ---- balance: pair heaviest with lightest, move the tail of the heavy list to the light warp
    std::vector<int> byc(NW);
    std::iota(byc.begin(), byc.end(), 0);
    std::sort(byc.begin(), byc.end(), [&](int a, int b) { return cost[a] > cost[b]; });
    std::vector<int> give_to(NW, -1), split_at(NW, -1);
    for (int i = 0; i < NW / 2; ++i) {
        int h = byc[i], l = byc[NW - 1 - i];
        if (cost[h] <= cost[l]) continue;
        int64_t move = (cost[h] - cost[l]) / 2, acc = 0;
        int k = (int)lists[h].size();
        while (k > 0) {
            int t = taylor_fma(s.jm[lists[h][k - 1]], J);
            if (acc + t > move) break;
            acc += t; --k;
        }
        if (k < (int)lists[h].size()) { give_to[h] = l; split_at[h] = k; }
    }
    s.pieces.assign(NW * 2, Piece{0, 0, 0, 0});
    auto emit = [&](int c1) {
        int v1[4]; g.coords(c1, v1);
        s.entries.push_back(g.pack(v1));
        s.entries.push_back(((uint32_t)s.supp_rec[c1] << 8) | s.jm[c1]);
    };
    for (int w = 0; w < NW; ++w) {
        int kend = split_at[w] >= 0 ? split_at[w] : (int)lists[w].size();
        int b = (int)s.entries.size() / 2;
        for (int k = 0; k < kend; ++k) emit(lists[w][k]);
        s.pieces[w * 2 + 0] = Piece{w, b, (int)s.entries.size() / 2, 1};
    }
    for (int h = 0; h < NW; ++h) {
        if (give_to[h] < 0) continue;
        int l = give_to[h];
        int b = (int)s.entries.size() / 2;
        for (int k = split_at[h]; k < (int)lists[h].size(); ++k) emit(lists[h][k]);
        s.pieces[l * 2 + 1] = Piece{h, b, (int)s.entries.size() / 2, 2};
    }
    return s;
}

}  // namespace rgplan
