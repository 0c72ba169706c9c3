// CPU check (host-compiled with nvcc, no GPU): the rough-conductor event derivatives of csrc/pb_rc.cuh — local forward-mode
// duals over (p, sh_n, a, q, n_q) and over the BSDF parameters — against central finite differences of the event's scalar.
#include <cstdio>
#include <cmath>
#include <random>

#include "../../psdr_cuda_b200/csrc/pb_rc.cuh"

using namespace pb;

static std::mt19937 rng(7);
static float U(float a, float b) { return std::uniform_real_distribution<float>(a, b)(rng); }
static float3 unit(float3 v) { const float n = std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z); return f3(v.x / n, v.y / n, v.z / n); }

struct Case {
    rc::Tex t;
    float x[15];
    bool primary, light, use_mis;
    float2 disk;
    float3 s3;
    float p_other, inv_cnt;
    float3 gA, gB;
};

// the event's scalar with the detached geometric term of pdf0 / pdf1 (direct.cpp:93,147) frozen at `g_det`, which is what the
// reference's derivative differentiates
static double value(const Case &c, const rc::Tex &t, const float *x, float g_det) {
    return (double)rc::branch_value<1>(t, x, 100, c.primary, f3(0.f), c.light, c.disk, c.p_other, c.use_mis, c.inv_cnt, c.gA, c.gB, g_det).v;
}
static float geometric_term(const float *x) {
    const float dx = x[9] - x[0], dy = x[10] - x[1], dz = x[11] - x[2];
    const float r2 = dx * dx + dy * dy + dz * dz, r = std::sqrt(r2);
    return std::fabs(x[12] * dx + x[13] * dy + x[14] * dz) / r / r2;
}

int main() {
    int bad = 0, checked = 0;
    for (int it = 0; it < 400; ++it) {
        Case c;
        c.t.au = U(0.15f, 0.6f); c.t.av = U(0.15f, 0.6f);
        c.t.eta = f3(U(0.2f, 1.5f), U(0.2f, 1.5f), U(0.2f, 1.5f)); c.t.k = f3(U(1.5f, 4.f), U(1.5f, 4.f), U(1.5f, 4.f)); c.t.spec = f3(U(0.5f, 1.f), U(0.5f, 1.f), U(0.5f, 1.f));
        c.primary = (it % 3) == 0; c.light = (it % 2) == 1; c.use_mis = (it % 4) < 2;
        const float3 p = f3(U(-1, 1), U(-1, 1), U(-1, 1)), n = unit(f3(U(-.3f, .3f), U(-.3f, .3f), 1.f));
        // incoming and outgoing directions in the upper hemisphere of n
        const float3 din = unit(f3(U(-.6f, .6f), U(-.6f, .6f), -1.f)), dout = unit(f3(U(-.6f, .6f), U(-.6f, .6f), 1.f));
        const float3 a = c.primary ? din : f3(p.x - 3.f * din.x, p.y - 3.f * din.y, p.z - 3.f * din.z);
        const float tq = U(2.f, 5.f);
        const float3 q = f3(p.x + tq * dout.x, p.y + tq * dout.y, p.z + tq * dout.z), nq = unit(f3(U(-.3f, .3f), U(-.3f, .3f), -1.f));
        const float xs[15] = {p.x, p.y, p.z, n.x, n.y, n.z, a.x, a.y, a.z, q.x, q.y, q.z, nq.x, nq.y, nq.z};
        for (int k = 0; k < 15; ++k) c.x[k] = xs[k];
        c.s3 = f3(U(0.05f, .95f), U(0.05f, .95f), U(0.f, 1.f));
        c.disk = square_to_uniform_disk_concentric(c.s3.x, c.s3.y);   // roughconductor.cpp:87 uses head<2>(sample)
        c.p_other = U(0.05f, 0.5f); c.inv_cnt = 1.f;
        c.gA = f3(U(-1, 1), U(-1, 1), U(-1, 1)); c.gB = c.light ? f3(0.f) : f3(U(-1, 1), U(-1, 1), U(-1, 1));
        const float gdet = geometric_term(c.x);
        const double v0 = value(c, c.t, c.x, gdet);
        if (!(std::fabs(v0) > 1e-4) || !std::isfinite(v0)) continue;
        // geometry: three passes of Dual<5>
        rc::GeomGrad g;
        if (!rc::branch_geom_grad(c.t, p, n, a, q, nq, c.primary, f3(0.f), c.light, c.disk, c.p_other, c.use_mis, c.inv_cnt, c.gA, c.gB, g)) continue;
        const float an[15] = {g.p.x, g.p.y, g.p.z, g.shn.x, g.shn.y, g.shn.z, g.a.x, g.a.y, g.a.z, g.q.x, g.q.y, g.q.z, g.nq.x, g.nq.y, g.nq.z};
        double scale = 1e-6;
        for (int k = 0; k < 15; ++k) scale = std::fmax(scale, std::fabs(an[k]));
        for (int k = 0; k < 15; ++k) {
            float xp[15], xm[15];
            for (int j = 0; j < 15; ++j) xp[j] = xm[j] = c.x[j];
            const float h = 2e-3f;
            xp[k] += h; xm[k] -= h;
            const double fd = (value(c, c.t, xp, gdet) - value(c, c.t, xm, gdet)) / (double)(xp[k] - xm[k]);
            ++checked;
            if (std::fabs(fd - an[k]) > 3e-2 * scale + 2e-3 * std::fabs(v0)) {
                if (bad < 10) std::printf("geometry mismatch case %d input %d: dual %g fd %g (value %g, primary %d light %d mis %d)\n", it, k, an[k], fd, v0, c.primary, c.light, c.use_mis);
                ++bad;
            }
        }
        // BSDF parameters: Dual<2> over (alpha_u, alpha_v), Fresnel per channel over (eta, k)
        {
            const Frame fr(n);
            const float3 wi = c.primary ? fr.to_local(f3(-a.x, -a.y, -a.z)) : fr.to_local(unit(f3(a.x - p.x, a.y - p.y, a.z - p.z)));
            const float3 dv = f3(q.x - p.x, q.y - p.y, q.z - p.z);
            const float r2 = dv.x * dv.x + dv.y * dv.y + dv.z * dv.z;
            const float3 wo = unit(dv), wo_l = fr.to_local(wo);
            const float G = std::fabs(nq.x * wo.x + nq.y * wo.y + nq.z * wo.z) / r2;
            rc::TexGrad tg;
            if (c.light) rc::light_branch_tex_grad(c.t, wi, wo_l, G, c.p_other, c.use_mis, c.inv_cnt, c.gA, tg);
            else rc::bsdf_branch_tex_grad(c.t, wi, wo_l, c.s3, G, c.p_other, c.use_mis, c.inv_cnt, c.gA, c.gB, tg);
            const float ga[11] = {tg.au, tg.av, tg.eta.x, tg.eta.y, tg.eta.z, tg.k.x, tg.k.y, tg.k.z, tg.spec.x, tg.spec.y, tg.spec.z};
            double sc = 1e-6;
            for (int k = 0; k < 11; ++k) sc = std::fmax(sc, std::fabs(ga[k]));
            for (int k = 0; k < 11; ++k) {
                rc::Tex tp = c.t, tm = c.t;
                float *fp = k == 0 ? &tp.au : k == 1 ? &tp.av : k < 5 ? (&tp.eta.x + (k - 2)) : k < 8 ? (&tp.k.x + (k - 5)) : (&tp.spec.x + (k - 8));
                float *fm = k == 0 ? &tm.au : k == 1 ? &tm.av : k < 5 ? (&tm.eta.x + (k - 2)) : k < 8 ? (&tm.k.x + (k - 5)) : (&tm.spec.x + (k - 8));
                const float h = 2e-3f;
                *fp += h; *fm -= h;
                const double fd = (value(c, tp, c.x, gdet) - value(c, tm, c.x, gdet)) / (double)(*fp - *fm);
                ++checked;
                if (std::fabs(fd - ga[k]) > 3e-2 * sc + 2e-3 * std::fabs(v0)) {
                    if (bad < 10) std::printf("parameter mismatch case %d param %d: dual %g fd %g (value %g)\n", it, k, ga[k], fd, v0);
                    ++bad;
                }
            }
        }
    }
    std::printf("rc_dual_check: %d comparisons, %d mismatches\n", checked, bad);
    if (bad == 0 && checked > 1000) std::printf("rc_dual_check: ok\n");
    return bad == 0 ? 0 : 1;
}
