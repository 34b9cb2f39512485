/* CPU ORACLE (test infrastructure only) — C restatement of noble-curves' `pippenger` for BLS12-381 G1.
 *
 * NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
 * legs may build, load or run this file.  It exists (a) as a faster checker than oracle/noble_ref.py at
 * 2^16..2^20 terms and (b) as the CPU baseline timed beside the GPU numbers.
 *
 * It follows the reference ALGORITHM line by line:
 *   /root/reference/src/abstract/curve.ts:863-905   pippenger: window rule :879-883, unsigned windows
 *                                                   MSB->LSB :888, bucket adds (bucket 0 included) :890-894,
 *                                                   running-sum reduction :897-900, c doublings :902
 *   /root/reference/src/abstract/weierstrass.ts:834-880  Point.add  (RCB alg. 1, a = 0, b3 = 3b = 12)
 *   /root/reference/src/abstract/weierstrass.ts:793-828  Point.double (RCB alg. 3, a = 0)
 *   /root/reference/src/abstract/weierstrass.ts:951-969  toAffine (ZERO -> (0,0))
 * The field arithmetic is the one deliberate difference: the reference computes (a*b) % p on BigInt
 * (modular.ts:50-54,956); here the same residues are carried in Montgomery form on six 64-bit limbs —
 * values after from-Montgomery are identical.  The reference is single-threaded; windows are
 * independent, so this port can also spread them over host threads (`threads` argument) to give the
 * strongest CPU baseline the algorithm allows.
 *
 * Parity status: pinned through oracle/noble_ref.py (itself pinned to the reference's golden vectors):
 * tests/test_oracle_c.py requires bit-identical affine results on the reference's own MSM test
 * constructions.
 *
 * Build: gcc -O3 -march=native -shared -fPIC -pthread oracle/ref_msm.c -o oracle/_build/libref_msm.so
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef uint64_t fp[6];

static const fp P = {0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                     0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL};
static const uint64_t PINV = 0x89f3fffcfffcfffdULL; /* -p^-1 mod 2^64 */
static fp R2;   /* 2^768 mod p, filled by init() */
static fp ONE;  /* 2^384 mod p */
static fp B3;   /* 12 in Montgomery form */
static int inited = 0;

static inline int fp_is0(const fp a) { return (a[0] | a[1] | a[2] | a[3] | a[4] | a[5]) == 0; }
static inline int fp_eq(const fp a, const fp b) {
  uint64_t t = 0;
  for (int i = 0; i < 6; i++) t |= a[i] ^ b[i];
  return t == 0;
}
static inline void fp_set(fp r, const fp a) { memcpy(r, a, sizeof(fp)); }

/* r = a - p if a >= p (a < 2p, `top` = carry limb) */
static inline void fp_reduce(fp r, uint64_t top) {
  fp d;
  u128 b = 0;
  for (int i = 0; i < 6; i++) {
    u128 t = (u128)r[i] - P[i] - (uint64_t)b;
    d[i] = (uint64_t)t;
    b = (t >> 64) & 1;
  }
  if (top || !b) fp_set(r, d);
}
static inline void fp_add(fp r, const fp a, const fp b) {
  u128 c = 0;
  for (int i = 0; i < 6; i++) {
    c += (u128)a[i] + b[i];
    r[i] = (uint64_t)c;
    c >>= 64;
  }
  fp_reduce(r, (uint64_t)c);
}
static inline void fp_sub(fp r, const fp a, const fp b) {
  u128 br = 0;
  for (int i = 0; i < 6; i++) {
    u128 t = (u128)a[i] - b[i] - (uint64_t)br;
    r[i] = (uint64_t)t;
    br = (t >> 64) & 1;
  }
  if (br) {
    u128 c = 0;
    for (int i = 0; i < 6; i++) {
      c += (u128)r[i] + P[i];
      r[i] = (uint64_t)c;
      c >>= 64;
    }
  }
}
/* Montgomery product (CIOS) */
static void fp_mul(fp r, const fp a, const fp b) {
  uint64_t t[8] = {0};
  for (int i = 0; i < 6; i++) {
    u128 c = 0;
    for (int j = 0; j < 6; j++) {
      c += (u128)a[j] * b[i] + t[j];
      t[j] = (uint64_t)c;
      c >>= 64;
    }
    c += t[6];
    t[6] = (uint64_t)c;
    t[7] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * PINV;
    c = (u128)m * P[0] + t[0];
    c >>= 64;
    for (int j = 1; j < 6; j++) {
      c += (u128)m * P[j] + t[j];
      t[j - 1] = (uint64_t)c;
      c >>= 64;
    }
    c += t[6];
    t[5] = (uint64_t)c;
    t[6] = t[7] + (uint64_t)(c >> 64);
  }
  memcpy(r, t, sizeof(fp));
  fp_reduce(r, t[6]);
}
static void fp_from_canon(fp r, const uint64_t* x) { fp_mul(r, x, R2); }
static void fp_to_canon(uint64_t* x, const fp a) {
  fp one = {1, 0, 0, 0, 0, 0};
  fp_mul(x, a, one);
}
static void fp_inv(fp r, const fp a) { /* a^(p-2) */
  fp e, acc, base;
  fp_set(e, P);
  e[0] -= 2;
  fp_set(acc, ONE);
  fp_set(base, a);
  for (int i = 0; i < 6; i++)
    for (int b = 0; b < 64; b++) {
      if ((e[i] >> b) & 1) fp_mul(acc, acc, base);
      fp_mul(base, base, base);
    }
  fp_set(r, acc);
}

typedef struct { fp X, Y, Z; } pt;

static void pt_zero(pt* p) { /* weierstrass.ts:687 ZERO = (0, 1, 0) */
  memset(p, 0, sizeof(*p));
  fp_set(p->Y, ONE);
}

/* weierstrass.ts:834-880, a = 0 (mulA == 0) */
static void pt_add(pt* r, const pt* p, const pt* q) {
  fp t0, t1, t2, t3, t4, t5, X3, Y3, Z3;
  fp_mul(t0, p->X, q->X);
  fp_mul(t1, p->Y, q->Y);
  fp_mul(t2, p->Z, q->Z);
  fp_add(t3, p->X, p->Y);
  fp_add(t4, q->X, q->Y);
  fp_mul(t3, t3, t4);
  fp_add(t4, t0, t1);
  fp_sub(t3, t3, t4);
  fp_add(t4, p->X, p->Z);
  fp_add(t5, q->X, q->Z);
  fp_mul(t4, t4, t5);
  fp_add(t5, t0, t2);
  fp_sub(t4, t4, t5);
  fp_add(t5, p->Y, p->Z);
  fp_add(X3, q->Y, q->Z);
  fp_mul(t5, t5, X3);
  fp_add(X3, t1, t2);
  fp_sub(t5, t5, X3);
  /* Z3 = mulA(t4) = 0 */
  fp_mul(X3, B3, t2);
  fp_set(Z3, X3);          /* Z3 = X3 + 0 */
  fp_sub(X3, t1, Z3);
  fp_add(Z3, t1, Z3);
  fp_mul(Y3, X3, Z3);
  fp_add(t1, t0, t0);
  fp_add(t1, t1, t0);
  /* t2 = mulA(t2) = 0 */
  fp_mul(t4, B3, t4);
  /* t1 = t1 + 0 ; t2 = mulA(t0 - 0) = 0 ; t4 = t4 + 0 */
  fp_mul(t0, t1, t4);
  fp_add(Y3, Y3, t0);
  fp_mul(t0, t5, t4);
  fp_mul(X3, t3, X3);
  fp_sub(X3, X3, t0);
  fp_mul(t0, t3, t1);
  fp_mul(Z3, t5, Z3);
  fp_add(Z3, Z3, t0);
  fp_set(r->X, X3);
  fp_set(r->Y, Y3);
  fp_set(r->Z, Z3);
}

/* weierstrass.ts:793-828, a = 0 */
static void pt_double(pt* r, const pt* p) {
  fp t0, t1, t2, t3, X3, Y3, Z3;
  fp_mul(t0, p->X, p->X);
  fp_mul(t1, p->Y, p->Y);
  fp_mul(t2, p->Z, p->Z);
  fp_mul(t3, p->X, p->Y);
  fp_add(t3, t3, t3);
  fp_mul(Z3, p->X, p->Z);
  fp_add(Z3, Z3, Z3);
  /* X3 = mulA(Z3) = 0 */
  fp_mul(Y3, B3, t2);
  /* Y3 = X3 + Y3 = Y3 */
  fp_sub(X3, t1, Y3);
  fp_add(Y3, t1, Y3);
  fp_mul(Y3, X3, Y3);
  fp_mul(X3, t3, X3);
  fp_mul(Z3, B3, Z3);
  /* t2 = mulA(t2) = 0; t3 = t0 - 0 ; t3 = mulA(t3) = 0 ; t3 = 0 + Z3 */
  fp_set(t3, Z3);
  fp_add(Z3, t0, t0);
  fp_add(t0, Z3, t0);
  /* t0 = t0 + t2(=0) */
  fp_mul(t0, t0, t3);
  fp_add(Y3, Y3, t0);
  fp_mul(t2, p->Y, p->Z);
  fp_add(t2, t2, t2);
  fp_mul(t0, t2, t3);
  fp_sub(X3, X3, t0);
  fp_mul(Z3, t2, t1);
  fp_add(Z3, Z3, Z3);
  fp_add(Z3, Z3, Z3);
  fp_set(r->X, X3);
  fp_set(r->Y, Y3);
  fp_set(r->Z, Z3);
}

static void init(void) {
  if (inited) return;
  /* ONE = 2^384 mod p by repeated doubling of 1; R2 = 2^768 mod p */
  fp x = {1, 0, 0, 0, 0, 0};
  for (int i = 0; i < 384; i++) fp_add(x, x, x);
  fp_set(ONE, x);
  for (int i = 0; i < 384; i++) fp_add(x, x, x);
  fp_set(R2, x);
  uint64_t twelve[6] = {12, 0, 0, 0, 0, 0};
  fp_from_canon(B3, twelve);
  inited = 1;
}

/* curve.ts:879-883 */
static int window_size(uint64_t n) {
  int wbits = 0;
  while (n) { wbits++; n >>= 1; }
  if (wbits > 12) return wbits - 3;
  if (wbits > 4) return wbits - 2;
  if (wbits > 0) return 2;
  return 1;
}

typedef struct {
  const pt* pts;
  const uint64_t* scalars; /* n x 4 limbs LE */
  uint64_t n;
  int c;
  int first_window, stride, num_windows;
  pt* window_res; /* resI per window index (position i / c) */
} job;

static uint32_t get_bits(const uint64_t* s, int off, int c) {
  int w = off >> 6, sh = off & 63;
  if (w >= 4) return 0;
  u128 v = s[w];
  if (w + 1 < 4) v |= (u128)s[w + 1] << 64;
  return (uint32_t)((v >> sh) & (((uint64_t)1 << c) - 1));
}

/* One window: curve.ts:889-900 */
static void do_window(const job* j, int widx, pt* buckets) {
  const uint64_t nb = (uint64_t)1 << j->c;
  for (uint64_t b = 0; b < nb; b++) pt_zero(&buckets[b]);
  const int off = widx * j->c;
  for (uint64_t k = 0; k < j->n; k++) {
    uint32_t wb = get_bits(j->scalars + 4 * k, off, j->c);
    pt_add(&buckets[wb], &buckets[wb], &j->pts[k]); /* bucket 0 is accumulated too (curve.ts:893) */
  }
  pt resI, sumI;
  pt_zero(&resI);
  pt_zero(&sumI);
  for (uint64_t b = nb - 1; b > 0; b--) {
    pt_add(&sumI, &sumI, &buckets[b]);
    pt_add(&resI, &resI, &sumI);
  }
  j->window_res[widx] = resI;
}

static void* worker(void* arg) {
  job* j = (job*)arg;
  pt* buckets = (pt*)malloc(sizeof(pt) << j->c);
  for (int w = j->first_window; w < j->num_windows; w += j->stride) do_window(j, w, buckets);
  free(buckets);
  return NULL;
}

/* pts: n x (x, y) canonical little-endian, 48 bytes per coordinate; (0,0) = ZERO (weierstrass.ts:716).
 * scalars: n x 32 bytes LE.  out_xy: 96 bytes canonical affine; returns is_inf.  fn_bits = Fn.BITS (255). */
int ref_pippenger_bls12_381_g1(const uint8_t* pts_b, const uint8_t* scalars_b, uint64_t n, int threads,
                               uint8_t* out_xy) {
  init();
  memset(out_xy, 0, 96);
  if (n == 0) return 1;
  pt* pts = (pt*)malloc(sizeof(pt) * n);
  for (uint64_t i = 0; i < n; i++) {
    uint64_t x[6], y[6];
    memcpy(x, pts_b + i * 96, 48);
    memcpy(y, pts_b + i * 96 + 48, 48);
    if (fp_is0(x) && fp_is0(y)) {
      pt_zero(&pts[i]);
    } else {
      fp_from_canon(pts[i].X, x);
      fp_from_canon(pts[i].Y, y);
      fp_set(pts[i].Z, ONE);
    }
  }
  const int fn_bits = 255;
  const int c = window_size(n);
  const int last_bits = ((fn_bits - 1) / c) * c; /* curve.ts:886 */
  const int num_windows = last_bits / c + 1;
  pt* wres = (pt*)malloc(sizeof(pt) * num_windows);
  if (threads < 1) threads = 1;
  if (threads > num_windows) threads = num_windows;
  job* jobs = (job*)malloc(sizeof(job) * threads);
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
  for (int t = 0; t < threads; t++) {
    jobs[t] = (job){pts, (const uint64_t*)scalars_b, n, c, t, threads, num_windows, wres};
    pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
  /* curve.ts:888,901-902: MSB -> LSB, sum = sum + resI; c doublings unless i == 0 */
  pt sum;
  pt_zero(&sum);
  for (int w = num_windows - 1; w >= 0; w--) {
    pt_add(&sum, &sum, &wres[w]);
    if (w != 0)
      for (int k = 0; k < c; k++) pt_double(&sum, &sum);
  }
  free(th);
  free(jobs);
  free(wres);
  free(pts);
  /* toAffine: weierstrass.ts:951-969 */
  if (fp_is0(sum.Z)) return 1;
  fp iz, x, y;
  fp_inv(iz, sum.Z);
  fp_mul(x, sum.X, iz);
  fp_mul(y, sum.Y, iz);
  uint64_t cx[6], cy[6];
  fp_to_canon(cx, x);
  fp_to_canon(cy, y);
  memcpy(out_xy, cx, 48);
  memcpy(out_xy + 48, cy, 48);
  return 0;
}

/* point-add count of the reference algorithm at size n: used to label extrapolations */
uint64_t ref_pippenger_add_count(uint64_t n, int fn_bits) {
  int c = window_size(n);
  int windows = ((fn_bits - 1) / c) + 1;
  return (uint64_t)windows * (n + 2 * (((uint64_t)1 << c) - 1)) + windows;
}

/* Synthetic inputs of test/slow-curves.test.ts:204-222: P_i = P_0 + i*S (affine in, affine out),
 * produced with the reference's own complete addition and one batched inversion (modular.ts:734-760). */
void ref_make_points_bls12_381_g1(const uint8_t* start_xy, const uint8_t* step_xy, uint64_t n, uint8_t* out) {
  init();
  pt cur, step;
  uint64_t t[6];
  memcpy(t, start_xy, 48); fp_from_canon(cur.X, t);
  memcpy(t, start_xy + 48, 48); fp_from_canon(cur.Y, t);
  fp_set(cur.Z, ONE);
  memcpy(t, step_xy, 48); fp_from_canon(step.X, t);
  memcpy(t, step_xy + 48, 48); fp_from_canon(step.Y, t);
  fp_set(step.Z, ONE);
  pt* all = (pt*)malloc(sizeof(pt) * n);
  fp* pref = (fp*)malloc(sizeof(fp) * n);
  fp acc;
  fp_set(acc, ONE);
  for (uint64_t i = 0; i < n; i++) {
    all[i] = cur;
    fp_set(pref[i], acc);
    fp_mul(acc, acc, cur.Z);
    pt_add(&cur, &cur, &step);
  }
  fp inv;
  fp_inv(inv, acc);
  for (uint64_t i = n; i-- > 0;) {
    fp iz, x, y;
    fp_mul(iz, inv, pref[i]);
    fp_mul(inv, inv, all[i].Z);
    fp_mul(x, all[i].X, iz);
    fp_mul(y, all[i].Y, iz);
    uint64_t cx[6], cy[6];
    fp_to_canon(cx, x);
    fp_to_canon(cy, y);
    memcpy(out + i * 96, cx, 48);
    memcpy(out + i * 96 + 48, cy, 48);
  }
  free(pref);
  free(all);
}
