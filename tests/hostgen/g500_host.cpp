// TEST ONLY: the host side of the device generator's arithmetic (csrc/cb_gen500.cuh is host/device code) prints edges
// [first, first + count) of the Graph500 2.1 stream for comparison with the reference's own generator.
#include <cstdio>
#include <cstdlib>
#include "cb_gen500.cuh"
int main(int argc, char** argv) {
    const int lgN = argc > 1 ? atoi(argv[1]) : 10;
    const long long first = argc > 2 ? atoll(argv[2]) : 0, count = argc > 3 ? atoll(argv[3]) : 16;
    static g500::Tables t;
    g500::build_tables(0, 0, &t);
    for (long long q = 0; q < count; ++q) {
        const unsigned long long ei = (unsigned long long)(first + q);
        g500::State s = t.seed;
        for (int b = 0; b < 4; ++b) { const unsigned v = (unsigned)((ei >> (8 * b)) & 0xFF); if (v) s = g500::apply(t.edge[b][v], s); }
        uint64_t a, c;
        g500::one_edge(s, lgN, t.val0, t.val1, &a, &c);
        printf("%llu %llu\n", (unsigned long long)a, (unsigned long long)c);
    }
    return 0;
}
