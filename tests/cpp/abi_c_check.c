/* The boundary is a C ABI: this file is compiled as C99 (tests/test_abi_cpu.py) to prove include/cvsteer_c.h needs no
 * C++, and run on the GPU box (tests/test_cpp_dropin_gpu.py) as the smallest possible C caller. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cvsteer_c.h"

int main(void)
{
    float taps[9];
    cvs_g2* h = NULL;
    cvs_batch b;
    int ndev = 0, i;
    float img[16 * 16], theta[16 * 16];
    if (cvs_g2_make_taps(1, 4, 0.67f, taps) != CVS_OK || taps[4] != 1.0f) return 2; /* g2(0) = exp(0) */
    memset(&b, 0, sizeof(b));
    if (cvs_g2_run_batch_dev(NULL, &b, 1u, CVS_STEER_DOMINANT, 0.f, NULL, NULL, NULL) != CVS_ERR_INVALID_ARG) return 3;
    if (cvs_device_count(&ndev) != CVS_OK || ndev < 1) {
        printf("no device: %s\n", cvs_last_error());
        return 0; /* CPU box: the no-fallback error path is the expected outcome */
    }
    for (i = 0; i < 256; ++i) img[i] = (float)((i * 37) % 251);
    if (cvs_g2_create(&h, 0, 4, 0.67f) != CVS_OK) return 4;
    if (cvs_g2_setup_host(h, img, 16, 16, 16 * sizeof(float)) != CVS_OK) return 5;
    if (cvs_g2_get_plane_host(h, CVS_THETA, theta, 16 * sizeof(float)) != CVS_OK) return 6;
    for (i = 0; i < 256; ++i)
        if (!(theta[i] >= -1.5708f && theta[i] <= 1.5708f)) return 7;
    cvs_g2_destroy(h);
    printf("abi_c ok (%s)\n", cvs_version());
    return 0;
}
