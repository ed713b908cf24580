/* A C99 consumer of include/xbitinfo_b200.h: what a cgo / JNI / plain C binding would compile.
 * Exercises the entry points that need no GPU (version, argument validation, error strings). */
#include <stdio.h>
#include <string.h>

#include "xbitinfo_b200.h"

int main(void) {
    float x[16] = {0};
    double mi[32];
    int64_t shape[2] = {4, 4}, chunks[2] = {2, 0};
    if (xbi_abi_version() != XBI_ABI_VERSION) return 1;
    if (xbi_workspace_bytes() == 0) return 2;
    if (xbi_bitround_host(x, x, XBI_F32, 16, 24, 0) != XBI_ERR_ARG) return 3;
    if (strstr(xbi_last_error(), "keepbits") == NULL) return 4;
    if (xbi_bitinformation_host(x, 99, 1, 16, 1, 0u, 0, 0, 0.99, mi, NULL, 0) != XBI_ERR_ARG) return 5;
    if (xbi_count_pairs_chunked(NULL, XBI_F32, 2, shape, chunks, 1, 0u, 0, NULL, NULL) != XBI_ERR_ARG) return 6;
    if (xbi_mutual_information_batched(NULL, 32, 0, 0, 0.99, NULL, NULL) != XBI_OK) return 7;
    if (xbi_shuffle(NULL, NULL, 3, 8, XBI_SHUFFLE_BYTES, 0, NULL) != XBI_ERR_ARG) return 8;
    printf("abi consumer ok (ABI %d, workspace %zu bytes)\n", xbi_abi_version(), xbi_workspace_bytes());
    return 0;
}
