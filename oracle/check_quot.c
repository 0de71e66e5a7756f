/* TEST INFRASTRUCTURE (oracle/): exhaustive check of the division shortcut of the pack kernels (kivi_common.cuh quot_to_half).
 *
 * Claim: for every finite fp16 a >= 0 and every finite fp16 s > 0,
 *     fast(a, s) = fp16(a * RN32(1/s))            when the product is clear of an fp16 rounding boundary (13 dropped mantissa
 *                                                  bits within 4 ulp of 0x1000), not fp16-subnormal and not NaN,
 *                  fp16(RN32(a / s))              otherwise,
 * equals  fp16(RN32(a / s))  -- the reference's fp16 `data.div_(scale)` (quant/new_pack.py:240, ATen: fp32 divide, fp16 round).
 * This program evaluates both for ALL 31744 x 31743 pairs (plain C, IEEE fp32, no contraction) and counts mismatches and
 * slow-path hits.  Usage: check_quot [stride]   (stride over s, default 1 = exhaustive). */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef _Float16 h16;
static inline h16 hbits(uint16_t b) { h16 h; memcpy(&h, &b, 2); return h; }
static inline uint16_t bitsh(h16 h) { uint16_t b; memcpy(&b, &h, 2); return b; }
static inline uint32_t fbits(float f) { uint32_t b; memcpy(&b, &f, 4); return b; }

int main(int argc, char** argv)
{
    const int stride = argc > 1 ? atoi(argv[1]) : 1;
    long long mism = 0, slow = 0, total = 0;
    for (uint32_t sb = 1; sb < 0x7c00; sb += stride) {              /* every positive finite fp16 scale (subnormals included) */
        const float s = (float)hbits((uint16_t)sb);
        volatile float rv = 1.0f / s;                               /* RN32(1/s): __frcp_rn */
        const float r = rv;
        for (uint32_t ab = 0; ab < 0x7c00; ++ab) {                  /* every finite fp16 a >= 0 */
            const float a = (float)hbits((uint16_t)ab);
            volatile float qv = a * r;
            float q = qv;
            const uint32_t b = fbits(q);
            const int sure = ((uint32_t)((b & 0x1fffu) - 0x0ffcu) > 8u) && (q >= 6.103515625e-05f || q == 0.0f);
            volatile float ev = a / s;                              /* RN32(a/s): __fdiv_rn */
            const float exact = ev;
            if (!sure) { q = exact; ++slow; }
            const uint16_t got = bitsh((h16)q), exp = bitsh((h16)exact);
            if (got != exp) { if (mism < 5) fprintf(stderr, "mismatch a=%04x s=%04x got=%04x exp=%04x\n", ab, sb, got, exp); ++mism; }
            ++total;
        }
    }
    printf("%lld %lld %lld\n", total, mism, slow);
    return mism != 0;
}
