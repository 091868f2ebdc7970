/* A plain C99 consumer of include/ddcb200.h: proves the boundary is a C ABI (no C++ / torch types) and doubles as the
 * minimal "how a non-Python caller uses the library" example.  Host buffers in, host buffers out.
 *
 *   ddc_c_client <taps.csv-less mode>: generates T = 256 Hann-windowed sinc taps, a tone at fc + 3.3 MHz, runs the DDC with
 *   D = 16 at fc = 100 MHz through ddcb200_run_host_f32 and checks the output against a direct double-precision
 *   evaluation of the reference formula (cwg.py:31-36, ddc.py:66,98,119) on 64 outputs.  Prints "C ABI OK" on success.
 *   With "--no-gpu" it only checks the entry points that need no device (version, out_len, error path).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ddcb200.h"

#define T 256
#define D 16
#define N (1 << 18)

int main(int argc, char** argv) {
    const double fs = 1712e6, fc = 100e6, pi = 3.14159265358979323846;
    if (ddcb200_version() != DDCB200_VERSION) return 2;
    if (ddcb200_out_len(N, T, D) != (N - T) / D + 1) return 3;
    if (ddcb200_create(NULL, 0, NULL, 0, 0) != DDCB200_EINVAL || strlen(ddcb200_last_error()) == 0) return 4;
    if (argc > 1 && strcmp(argv[1], "--no-gpu") == 0) {
        printf("C ABI surface OK (no GPU calls)\n");
        return 0;
    }
    double taps[T], sum = 0.0;
    for (int i = 0; i < T; ++i) {
        const double t = (i - (T - 1) / 2.0) * 0.8 / D, w = 0.5 - 0.5 * cos(2 * pi * i / (T - 1));
        taps[i] = w * (fabs(t) < 1e-12 ? 1.0 : sin(pi * t) / (pi * t));
        sum += taps[i];
    }
    float* x = (float*)ddcb200_host_alloc(sizeof(float) * N);           /* pinned, like ddc_host_gpu.py:45-58 */
    const long long m = (long long)ddcb200_out_len(N, T, D);
    ddcb200_c64* y = (ddcb200_c64*)ddcb200_host_alloc(sizeof(ddcb200_c64) * (size_t)m);
    if (!x || !y) return 5;
    for (long long n = 0; n < N; ++n) x[n] = (float)floor(100.0 * cos(2 * pi * (fc + 3.3e6) / fs * (double)n) + 0.5);
    ddcb200_t* h = NULL;
    if (ddcb200_create(&h, 0, taps, T, D) != DDCB200_OK) {
        fprintf(stderr, "create: %s\n", ddcb200_last_error());
        return 6;
    }
    const double step = floor((double)N / (fs / fc)) / (double)(N - 1);   /* int(N fc / fs) / (N - 1), cwg.py:31-33 */
    if (ddcb200_run_host_f32(h, x, N, 1, N, step, 0, y, m) != DDCB200_OK) {
        fprintf(stderr, "run: %s\n", ddcb200_last_error());
        return 7;
    }
    double worst = 0.0, scale = 0.0;
    for (long long mm = 1000; mm < 1064; ++mm) {
        double re = 0.0, im = 0.0;
        for (int i = 0; i < T; ++i) {                                      /* ddc.py:98: sum_i taps[i] mix[mD + T-1-i] / sum */
            const long long n = mm * D + T - 1 - i;
            const double ph = fmod((double)n * step, 1.0);
            re += taps[i] * x[n] * cos(2 * pi * ph);
            im -= taps[i] * x[n] * sin(2 * pi * ph);
        }
        re /= sum;
        im /= sum;
        const double e = hypot(y[mm].re - re, y[mm].im - im), a = hypot(re, im);
        if (e > worst) worst = e;
        if (a > scale) scale = a;
    }
    printf("variant %s, launches %lld, worst |err| / max|y| = %.3e\n", ddcb200_last_variant(h), (long long)ddcb200_launch_count(h),
           worst / scale);
    ddcb200_destroy(h);
    ddcb200_host_free(x);
    ddcb200_host_free(y);
    if (!(worst <= 1e-5 * scale)) return 8;
    printf("C ABI OK\n");
    return 0;
}
