// Piece statistics of the convolution plan for any number of (virtual) warps -- host only, includes the planner.
// usage: plan_pieces <dim> <L> <J> <iminus.bin (uint8 [J+1][M] wrapped)> <half 0|1> <n_warps>
// prints one line per piece: stage box rank begin end cost   (cost: the planner's issue model, summed over the entries)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../rg_b200/csrc/rg_plan.hpp"
int main(int argc, char** argv) {
    if (argc < 7) return 2;
    const int dim = atoi(argv[1]), L = atoi(argv[2]), J = atoi(argv[3]), half = atoi(argv[5]), nw = atoi(argv[6]);
    rgplan::Geometry g; g.init(dim, L, J);
    std::vector<uint8_t> im((size_t)g.vecjl);
    FILE* f = fopen(argv[4], "rb");
    if (!f || fread(im.data(), 1, im.size(), f) != im.size()) return 3;
    fclose(f);
    std::vector<uint16_t> mask(g.M, 0);
    for (int c = 0; c < g.M; ++c) {
        const int w = g.wrapped_of_centred(c);
        for (int j = 0; j <= J; ++j) if (im[(size_t)j * g.M + w]) mask[c] |= (uint16_t)(1u << j);
    }
    rgplan::Schedule s = rgplan::make_schedule(g, mask, false, half != 0, nw);
    if (s.n_warps < 0) { printf("# plan failed\n"); return 1; }
    const int nbs = s.n_boxes / s.n_stages;
    printf("# warps %d stages %d boxes %d maxp %d cost_max %lld cost_total %lld\n", s.n_warps, s.n_stages, s.n_boxes, s.maxp,
           (long long)s.cost_max, (long long)s.cost_total);
    for (size_t i = 0; i < s.pieces.size(); ++i) {
        const rgplan::Piece& p = s.pieces[i];
        if (p.box < 0) continue;
        long long cost = 0;
        for (int e = p.begin; e < p.end; ++e) {
            const int jm = (int)(s.entries[2 * e] >> 24);
            cost += (long long)(2.65 * (double)rgplan::taylor_fma(rgplan::level_class(jm, J), J)) + 67;
        }
        printf("%d %d %d %d %d %lld\n", p.box / nbs, p.box, p.rank, p.begin, p.end, cost);
    }
    return 0;
}
