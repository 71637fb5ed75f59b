/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the workload bodies and of the reference's
 * result placement, used to check the CUDA path at BASELINE.json's full sizes (1e8 tasks) in
 * seconds.  Mirrors oracle/bodies.py function for function; both are pinned against the golden
 * vectors produced by the real reference pool (tests/golden/, tests/test_oracle.py).
 *
 * Build (done by __graft_entry__.build() / oracle/Makefile):
 *   gcc -O2 -ffp-contract=off -shared -fPIC oracle/fbr_oracle.c -o oracle/_build/libfbr_oracle.so
 * -ffp-contract=off matters: `x*x + y*y` must stay three separately rounded IEEE-754 double
 * operations, as CPython evaluates examples/pi_estimation.py:11.
 *
 * Nothing under fiber_b200/ links or loads this file.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---------------- Philox4x32-10 (Random123 philox.h; not in the reference) ------------------- */
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u
#define PI_KEY0 0xF1BE5EEDu
#define PI_KEY1 0x00000000u

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)PHILOX_M0 * c0, p1 = (uint64_t)PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += PHILOX_W0; k1 += PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* examples/pi_estimation.py:9-11 with random.random() replaced by a Philox block keyed by p. */
static inline uint8_t pi_inside_det(int64_t p) {
    uint32_t ctr[4] = {(uint32_t)((uint64_t)p), (uint32_t)((uint64_t)p >> 32), 0u, 0u};
    uint32_t key[2] = {PI_KEY0, PI_KEY1}, o[4];
    orc_philox4x32_10(ctr, key, o);
    double x = ((double)(o[0] >> 5) * 67108864.0 + (double)(o[1] >> 6)) / 9007199254740992.0;
    double y = ((double)(o[2] >> 5) * 67108864.0 + (double)(o[3] >> 6)) / 9007199254740992.0;
    double xx = x * x, yy = y * y;
    double s = xx + yy;
    return s < 1.0;
}

/* [pi_inside_det(start + i*step) for i in range(n)] -> out (may be NULL); returns the count. */
int64_t orc_pi_inside_range(int64_t start, int64_t step, int64_t n, uint8_t* out) {
    int64_t count = 0;
#pragma omp parallel for reduction(+ : count) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        uint8_t r = pi_inside_det((int64_t)((uint64_t)start + (uint64_t)i * (uint64_t)step));
        if (out) out[i] = r;
        count += r;
    }
    return count;
}

/* ---------------- examples/parzen_estimation.py:6-15 ------------------------------------------ */
/* k_n for one width, float64 arithmetic (the reference's). dims = point_x.shape[0]. */
int64_t orc_parzen_count_f64(const double* xs, int64_t n, int dims, const double* px, double h) {
    int64_t k = 0;
    for (int64_t i = 0; i < n; ++i) {
        int inside = 1;
        for (int d = 0; d < dims; ++d)
            if (fabs((px[d] - xs[i * dims + d]) / h) > 0.5) { inside = 0; break; }
        k += inside;
    }
    return k;
}

/* Same in float32 arithmetic: what the north-star's fp32 device path computes. */
int64_t orc_parzen_count_f32(const float* xs, int64_t n, int dims, const float* px, float h) {
    int64_t k = 0;
    for (int64_t i = 0; i < n; ++i) {
        int inside = 1;
        for (int d = 0; d < dims; ++d) {
            volatile float q = (px[d] - xs[i * dims + d]) / h; /* volatile: no excess precision */
            if (fabsf(q) > 0.5f) { inside = 0; break; }
        }
        k += inside;
    }
    return k;
}

/* (h, (k_n / n) / h**power): examples/parzen_estimation.py:15, power = point_x.shape[1]. */
double orc_parzen_density(int64_t k, int64_t n, double h, int power) {
    return ((double)k / (double)n) / pow(h, (double)power);
}

/* ---------------- synthetic 4 KB payload map (BASELINE.json config 4) ------------------------- */
#define PAYLOAD_WORDS 1024
#define PAYLOAD_SEED 0xF1BE5ull
#define PAYLOAD_MUL 2654435761u

uint64_t orc_splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

void orc_payload_records(int64_t t0, int64_t n, uint32_t* out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n * PAYLOAD_WORDS; ++i)
        out[i] = (uint32_t)orc_splitmix64(PAYLOAD_SEED ^ (uint64_t)(t0 * PAYLOAD_WORDS + i));
}

void orc_payload_map(int64_t t0, int64_t n, const uint32_t* in, uint32_t* out) {
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < n; ++t)
        for (int j = 0; j < PAYLOAD_WORDS; ++j)
            out[t * PAYLOAD_WORDS + j] = in[t * PAYLOAD_WORDS + j] * PAYLOAD_MUL + (uint32_t)(t0 + t);
}

void orc_payload_checksum(int64_t n, const uint32_t* in, uint32_t* out) {
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < n; ++t) {
        uint32_t s = 0;
        for (int j = 0; j < PAYLOAD_WORDS; ++j) s += in[t * PAYLOAD_WORDS + j];
        out[t] = s;
    }
}

/* ---------------- result placement (fiber/pool.py:666-679 Inventory.get) ---------------------- */
/* Messages (seq, batch, idx, result) arrive in arbitrary order; each is stored at inventory[idx].
 * Restated for fixed-size results: `order` is the arrival permutation of chunk numbers, chunk c
 * carries results for indices [c*chunk, min(n,(c+1)*chunk)) packed back to back in `ring`. */
void orc_place_by_index(const uint8_t* ring, const int64_t* order, int64_t n_chunks, int64_t chunk,
                        int64_t n, int64_t result_bytes, uint8_t* out) {
    int64_t off = 0;
    for (int64_t a = 0; a < n_chunks; ++a) {
        int64_t c = order[a], first = c * chunk, cnt = (first + chunk <= n) ? chunk : (n - first);
        memcpy(out + first * result_bytes, ring + off, (size_t)(cnt * result_bytes));
        off += cnt * result_bytes;
    }
}

/* int64 bodies of tests/test_pool.py:18-25 (wrap-around like the device's checked i64 multiply
 * never triggers in the tested ranges; overflow is reported through *ovf). */
void orc_square_i64(const int64_t* x, int64_t n, int64_t* out, int* ovf) {
    for (int64_t i = 0; i < n; ++i) *ovf |= __builtin_mul_overflow(x[i], x[i], &out[i]);
}

/* Single-task entry so a Python callable can run the deterministic body at C speed inside the CPU
 * pool port's workers (keeps the CPU arm messaging-bound, like the reference's own is_inside). */
int orc_pi_inside_one(int64_t p) { return pi_inside_det(p); }
