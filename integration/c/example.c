/* C user of the reference's own two-function ABI (include/yume_pdq.h:120-147 of the reference), linked against
 * libyume_pdq_cuda.so instead of libyume_pdq: hash one frame, then visit its 7 dihedrals through the callback.
 * Plain C11 -- also the check that include/yume_pdq_cuda.h is a valid C header.
 * Build: gcc -std=c11 -Iinclude integration/c/example.c -Lyume_pdq_b200 -lyume_pdq_cuda -lm
 * Prints one line per hash: "<dihedral packed as 8 hex digits> <quality> <64 hex digits>". */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "yume_pdq_cuda.h"

struct visit_ctx {
    const uint8_t *hash; /* the library rewrites this buffer before every callback */
    int calls;
};

static bool on_dihedral(void *p, uint32_t dihedral, float threshold, float quality)
{
    struct visit_ctx *ctx = (struct visit_ctx *)p;
    (void)threshold;
    printf("%08x %.6f ", dihedral, quality);
    for (int i = 0; i < 32; i++)
        printf("%02x", ctx->hash[i]);
    printf("\n");
    ctx->calls++;
    return true;
}

int main(void)
{
    float *input = malloc(512 * 512 * sizeof(float));
    float *buf1 = calloc(128 * 128, sizeof(float));
    float tmp[128] = {0}, pdqf[256], threshold = 0.0f;
    uint8_t output[32];
    if (!input || !buf1)
        return 3;
    for (int y = 0; y < 512; y++)
        for (int x = 0; x < 512; x++)
            input[y * 512 + x] = (float)((x * 7 + y * 3 + (x * y) / 64) % 256);
    float quality = yume_pdq_hash_smart_kernel(input, &threshold, output, buf1, tmp, pdqf);
    if (isnan(quality)) {
        fprintf(stderr, "no usable sm_100 CUDA device (this library has no CPU fallback)\n");
        return 2;
    }
    printf("%08x %.6f ", YUME_PDQ_DIHEDRAL_NORMAL.packed, quality);
    for (int i = 0; i < 32; i++)
        printf("%02x", output[i]);
    printf("\n");
    struct visit_ctx ctx = {output, 0};
    if (!yume_pdq_visit_dihedrals_smart_kernel(&ctx, threshold, output, pdqf, on_dihedral) || ctx.calls != 7)
        return 1;
    printf("version %s\n", YUME_PDQ_VERSION_STRING);
    free(input);
    free(buf1);
    return 0;
}
