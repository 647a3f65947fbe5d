/* hvb200_mtjump.c — jump-ahead polynomials for MT19937 (host, computed once per process).
 *
 * The DZ2019-MC direction stream is the reference's MT19937 stream (c/mt19937/mt19937.c:18-55).  One CTA can
 * advance that recurrence by only ~2.4 words/ns (x[n] needs x[n-227]), which made the generator 18x more
 * expensive than the max-min kernel.  The state sequence x_0, x_1, ... is linear over GF(2): with phi its
 * characteristic polynomial (degree 19937) and g_J(t) = t^J mod phi(t),
 *
 *      x_{n+J} = XOR over { i : coefficient i of g_J is set } of x_{n+i}            for every n,
 *
 * so the state J words ahead is a GF(2) convolution of 20 560 consecutive state words with g_J (the "sliding
 * window" form of Haramoto, Matsumoto, Nishimura, Panneton, L'Ecuyer, "Efficient jump ahead for F2-linear
 * random number generators", 2008) — embarrassingly parallel on the GPU (hvb200_mcgen.cuh).  This file
 * supplies phi (Berlekamp-Massey on one output bit of a generic state) and the polynomials g_{J 2^k}.
 * Nothing here depends on the seed.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/hvapprox_b200.h"
#include "hvb200_internal.h"

#define MT_N 624
#define MT_M 397
#define DEG 19937
#define PW 313                 /* 64-bit words of a polynomial of degree <= 19968+63 */

static void mt_next_block(uint32_t *x)
{   /* in-place block update, c/mt19937/mt19937.c:18-55 */
    for (int i = 0; i < MT_N; i++) {
        const uint32_t y = (x[i] & 0x80000000u) | (x[(i + 1) % MT_N] & 0x7fffffffu);
        x[i] = x[(i + MT_M) % MT_N] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
}

static inline int getbit(const uint64_t *p, int i) { return (int)((p[i >> 6] >> (i & 63)) & 1u); }
static inline void flipbit(uint64_t *p, int i) { p[i >> 6] ^= (uint64_t)1 << (i & 63); }

/* dst ^= src << shift, both `words` long (bits shifted out at the top are dropped) */
static void xor_shifted(uint64_t *dst, const uint64_t *src, int shift, int words)
{
    const int ws = shift >> 6, bs = shift & 63;
    if (bs == 0) {
        for (int i = words - 1; i >= ws; i--) dst[i] ^= src[i - ws];
    } else {
        for (int i = words - 1; i > ws; i--) dst[i] ^= (src[i - ws] << bs) | (src[i - ws - 1] >> (64 - bs));
        if (ws < words) dst[ws] ^= src[0] << bs;
    }
}

static uint64_t g_phi[PW];                 /* monic, degree 19937 */
static uint64_t g_phi_sh[64][2 * PW];      /* phi << s, s = 0..63 (for the reduction) */
static int g_phi_ready = 0;

/* Berlekamp-Massey over GF(2) on bit 0 of the state words x_1, x_2, ... of a generic state */
static int compute_phi(void)
{
    const int NB = 2 * DEG + 64;
    uint32_t x[MT_N];
    uint32_t s = 19650218u;
    for (int i = 0; i < MT_N; i++) { x[i] = s; s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)i + 1u; }
    unsigned char *seq = malloc((size_t)NB);
    if (!seq) return -1;
    int n = 0;
    mt_next_block(x);                       /* skip the seed block: only bit 31 of its first word is state */
    while (n < NB) {
        mt_next_block(x);
        for (int i = 0; i < MT_N && n < NB; i++) seq[n++] = (unsigned char)(x[i] & 1u);
    }
    /* C, B: connection polynomials; R: reversed window, bit i = seq[n - i] */
    static uint64_t C[PW], B[PW], T[PW], R[PW];
    memset(C, 0, sizeof C); memset(B, 0, sizeof B); memset(R, 0, sizeof R);
    C[0] = 1; B[0] = 1;
    int L = 0, m = 1;
    for (n = 0; n < NB; n++) {
        /* R <<= 1; R_0 = seq[n] */
        for (int i = PW - 1; i > 0; i--) R[i] = (R[i] << 1) | (R[i - 1] >> 63);
        R[0] = (R[0] << 1) | seq[n];
        uint64_t acc = 0;
        const int lw = (L >> 6) + 1;
        for (int i = 0; i < lw && i < PW; i++) acc ^= C[i] & R[i];
        const int d = __builtin_parityll(acc);
        if (!d) { m++; continue; }
        if (2 * L <= n) {
            memcpy(T, C, sizeof C);
            xor_shifted(C, B, m, PW);
            L = n + 1 - L;
            memcpy(B, T, sizeof B);
            m = 1;
        } else {
            xor_shifted(C, B, m, PW);
            m++;
        }
        if (L > DEG) break;
    }
    free(seq);
    if (L != DEG) return -1;
    /* s_n = sum_{j=1..L} C_j s_{n-j}  <=>  sum_i phi_i s_{n+i} = 0 with phi_i = C_{L-i} */
    memset(g_phi, 0, sizeof g_phi);
    for (int i = 0; i <= DEG; i++) if (getbit(C, DEG - i)) flipbit(g_phi, i);
    if (!getbit(g_phi, DEG) || !getbit(g_phi, 0)) return -1;
    for (int sft = 0; sft < 64; sft++) {
        memset(g_phi_sh[sft], 0, sizeof g_phi_sh[sft]);
        xor_shifted(g_phi_sh[sft], g_phi, 0, PW);
        if (sft) {
            uint64_t tmp[2 * PW];
            memset(tmp, 0, sizeof tmp);
            xor_shifted(tmp, g_phi_sh[sft], sft, 2 * PW);   /* g_phi_sh[sft] currently holds phi in the low words */
            memcpy(g_phi_sh[sft], tmp, sizeof tmp);
        }
    }
    return 0;
}

/* a (degree < 2*19968) reduced modulo phi in place; the result occupies the low PW words */
static void reduce(uint64_t *a, int top_bit)
{
    for (int i = top_bit; i >= DEG; i--) {
        if (!((a[i >> 6] >> (i & 63)) & 1u)) continue;
        const int sh = i - DEG, ws = sh >> 6;
        const uint64_t *p = g_phi_sh[sh & 63];
        for (int w = 0; w < PW + 1 && ws + w < 2 * PW; w++) a[ws + w] ^= p[w];
    }
}
static uint64_t spread32(uint32_t v)
{   /* bit i -> bit 2i */
    uint64_t x = v;
    x = (x | (x << 16)) & 0x0000ffff0000ffffull;
    x = (x | (x << 8)) & 0x00ff00ff00ff00ffull;
    x = (x | (x << 4)) & 0x0f0f0f0f0f0f0f0full;
    x = (x | (x << 2)) & 0x3333333333333333ull;
    x = (x | (x << 1)) & 0x5555555555555555ull;
    return x;
}
static void poly_square(uint64_t *r /* PW */)
{
    uint64_t a[2 * PW];
    for (int i = 0; i < PW; i++) {
        a[2 * i] = spread32((uint32_t)r[i]);
        a[2 * i + 1] = spread32((uint32_t)(r[i] >> 32));
    }
    reduce(a, 2 * DEG);
    memcpy(r, a, PW * sizeof(uint64_t));
}
static void poly_mul_t(uint64_t *r /* PW */)
{   /* r = r * t mod phi */
    uint64_t a[2 * PW];
    memset(a, 0, sizeof a);
    for (int i = PW - 1; i > 0; i--) a[i] = (r[i] << 1) | (r[i - 1] >> 63);
    a[0] = r[0] << 1;
    a[PW] = r[PW - 1] >> 63;
    reduce(a, DEG + 1);
    memcpy(r, a, PW * sizeof(uint64_t));
}

static pthread_mutex_t g_mtx = PTHREAD_MUTEX_INITIALIZER;

/* t^J mod phi as 624 32-bit words (bit i of word w = coefficient 32 w + i).  Returns 0 on success. */
int hvb200_mtjump_poly(uint64_t J, uint32_t out[MT_N])
{
    pthread_mutex_lock(&g_mtx);
    if (!g_phi_ready) {
        if (compute_phi()) { pthread_mutex_unlock(&g_mtx); hvb200_set_error("MT19937 characteristic polynomial: Berlekamp-Massey failed"); return -1; }
        g_phi_ready = 1;
    }
    pthread_mutex_unlock(&g_mtx);
    uint64_t r[PW];
    memset(r, 0, sizeof r);
    r[0] = 1;
    for (int b = 63; b >= 0; b--) {
        poly_square(r);
        if ((J >> b) & 1u) poly_mul_t(r);
    }
    for (int w = 0; w < MT_N; w++) out[w] = (uint32_t)(r[w >> 1] >> (32 * (w & 1)));
    return 0;
}

/* the polynomial of twice the jump: out = in^2 mod phi */
int hvb200_mtjump_double(const uint32_t in[MT_N], uint32_t out[MT_N])
{
    if (!g_phi_ready) { hvb200_set_error("internal: mtjump not initialised"); return -1; }
    uint64_t r[PW];
    memset(r, 0, sizeof r);
    for (int w = 0; w < MT_N; w++) r[w >> 1] |= (uint64_t)in[w] << (32 * (w & 1));
    poly_square(r);
    for (int w = 0; w < MT_N; w++) out[w] = (uint32_t)(r[w >> 1] >> (32 * (w & 1)));
    return 0;
}

/* Host restatement of the device jump (tests, and the reference for the device kernels): state_out = the
   624-word window of the state sequence that starts J words after state_in's. */
int hvb200_mtjump_apply_host(const uint32_t state_in[MT_N], const uint32_t poly[MT_N], uint32_t state_out[MT_N])
{
    const int NBLK = 33;                               /* 624 + 33*624 = 21216 >= 19937 + 623 words */
    uint32_t *seq = malloc((size_t)(NBLK + 1) * MT_N * sizeof(uint32_t));
    if (!seq) { hvb200_set_error("out of memory"); return -1; }
    memcpy(seq, state_in, MT_N * sizeof(uint32_t));
    uint32_t x[MT_N];
    memcpy(x, state_in, sizeof x);
    for (int b = 1; b <= NBLK; b++) {
        mt_next_block(x);
        memcpy(seq + (size_t)b * MT_N, x, sizeof x);
    }
    memset(state_out, 0, MT_N * sizeof(uint32_t));
    for (int i = 0; i < DEG; i++) {
        if (!((poly[i >> 5] >> (i & 31)) & 1u)) continue;
        for (int p = 0; p < MT_N; p++) state_out[p] ^= seq[i + p];
    }
    free(seq);
    return 0;
}
