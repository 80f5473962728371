/* Plain C host for the C-ABI (include/caviar_b200.h): the compute_distances_parallel operator of CAVIAR-1.0.py:110-122
 * without any Python.  Build (the library is built by `python -c "import __graft_entry__ as g; g.build()"`):
 *
 *   gcc -O2 -Iinclude examples/pair_distances.c -o pair_distances \
 *       -Lcaviar_b200/csrc -lcaviar_b200 -Wl,-rpath,$PWD/caviar_b200/csrc
 *
 * On a machine without a B200 it prints the library's refusal (there is no CPU fallback) and exits with status 2. */
#include <stdio.h>
#include <stdlib.h>

#include "caviar_b200.h"

int main(void) {
    enum { T = 4, N = 5, P = 3 };
    float xyz[T][N][3];
    const int32_t pairs[P][2] = {{0, 1}, {0, 4}, {2, 3}};
    float dist[T][P];
    for (int t = 0; t < T; ++t)
        for (int a = 0; a < N; ++a) {
            xyz[t][a][0] = 3.0f * a + 0.1f * t;
            xyz[t][a][1] = 4.0f * a;
            xyz[t][a][2] = 0.0f;
        }
    printf("caviar_b200 ABI version %d\n", caviar_abi_version());
    const int rc = caviar_pair_distances_host(&xyz[0][0][0], T, N, &pairs[0][0], P, &dist[0][0]);
    if (rc != CV_OK) {
        printf("caviar_pair_distances_host: error %d: %s\n", rc, caviar_last_error());
        return rc == CV_E_NO_DEVICE ? 2 : 1;
    }
    for (int t = 0; t < T; ++t) {
        printf("frame %d:", t + 1);
        for (int p = 0; p < P; ++p) printf(" %.4f", dist[t][p]);
        printf("\n");
    }
    caviar_release_cache();
    return 0;
}
