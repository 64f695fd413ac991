/* CPU replay of csrc/stepgen.cu's make_steps for cascade entries: which steps of bench.py's config-4 leg come out at
 * infinity (gamma_distributed at ry == 1)?  Built and driven by tools/replay_config4_stepgen.py.  Diagnostic only. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>

typedef struct { uint64_t x; uint32_t a; uint64_t draws, zero_draws; } Mwc;

static double co(Mwc *r)
{
    r->x = (r->x & 0xffffffffull) * r->a + (r->x >> 32);
    r->draws++;
    if ((uint32_t)r->x == 0u) r->zero_draws++;
    return (double)(uint32_t)(r->x & 0xffffffffull) / 4294967296.0;
}
static double oc(Mwc *r) { return 1.0 - co(r); }

static double gamma_distributed(double shape, Mwc *rng)
{
    double x;
    if (shape < 1.) {
        const double c = 1. / shape, d = (1. - shape) * pow(shape, shape / (1. - shape));
        double z, e;
        do { z = -log(oc(rng)); e = -log(oc(rng)); x = pow(z, c); } while (z + e < d + x);
    } else {
        const double b = shape - log(4.0), l = sqrt(2. * shape - 1.0), cheng = 1.0 + log(4.5);
        float y, z, r;
        do {
            const double rx = oc(rng), ry = oc(rng);
            y = (float)(log(ry / (1. - ry)) / l);
            x = shape * (double)(float)exp((double)y);
            z = (float)(rx * ry * ry);
            r = (float)(b + (shape + l) * (double)y - x);
        } while ((double)r < 4.5 * (double)z - cheng && (double)r < (double)(float)log((double)z));
    }
    return x;
}

/* one launch: `total` steps of ONE cascade entry (pa, pb); thread me takes steps me, me + threads, ... */
long replay_launch(uint64_t *x, const uint32_t *a, uint32_t threads, uint64_t total, double pa, double pb, uint64_t *draws, uint64_t *zero_draws,
                   long long *first_bad_step, uint32_t *first_bad_thread)
{
    long bad = 0;
    for (uint32_t me = 0; me < threads; ++me) {
        Mwc rng = {x[me], a[me], 0, 0};
        for (uint64_t j = me; j < total; j += threads) {
            const double along = pb * gamma_distributed(pa, &rng);
            co(&rng);   /* cos_val */
            co(&rng);   /* random_value */
            if (!isfinite(along)) {
                if (bad == 0) { *first_bad_step = (long long)j; *first_bad_thread = me; }
                ++bad;
            }
        }
        x[me] = rng.x;
        *draws += rng.draws;
        *zero_draws += rng.zero_draws;
    }
    return bad;
}
