#include <cstdio>
#include <cstdlib>
#include "fe29.h"
// prints a, b, a*b, a^2 as limb lists for a python checker
int main() {
    srand(7);
    for (int it = 0; it < 2000; it++) {
        fe9 a, b;
        for (int i = 0; i < 9; i++) {
            uint32_t lim = (i == 8) ? (1u << 24) : (1u << 29);
            uint32_t slack = (it % 3 == 0) ? 0 : (it % 3 == 1 ? (1u << 18) : lim);   // magnitude 1, 1+, 2
            uint64_t ra = ((uint64_t)rand() << 20) ^ rand(), rb = ((uint64_t)rand() << 20) ^ rand();
            a.v[i] = (uint32_t)(ra % (lim + slack)); b.v[i] = (uint32_t)(rb % (lim + slack));
            if (it % 7 == 0) { a.v[i] = lim + slack - 1; b.v[i] = lim + slack - 1; }
        }
        fe9 m = fe9_mul(a, b), s = fe9_sqr(a);
        for (int i = 0; i < 9; i++) printf("%u ", a.v[i]);
        for (int i = 0; i < 9; i++) printf("%u ", b.v[i]);
        for (int i = 0; i < 9; i++) printf("%u ", m.v[i]);
        for (int i = 0; i < 9; i++) printf("%u ", s.v[i]);
        printf("\n");
    }
}
