// CPU check of the reverse-mode building blocks in csrc/pb_adjoint_math.cuh against central finite differences.
// Built and run by tests/test_adjoint_math.py (nvcc host compilation only; no GPU needed).
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <vector>

#include "../../psdr_cuda_b200/csrc/pb_adjoint_math.cuh"

using namespace pb;

static double rnd() { return (double)rand() / RAND_MAX * 2.0 - 1.0; }
static int fails = 0;

// compare analytic gradient `g` of scalar function F at x with central differences evaluated in double
static void check(const char *name, std::function<double(const std::vector<double> &)> F, const std::vector<double> &x, const std::vector<double> &g) {
    for (size_t i = 0; i < x.size(); ++i) {
        std::vector<double> a = x, b = x;
        const double h = 1e-4 * (1.0 + fabs(x[i]));
        a[i] += h; b[i] -= h;
        const double fd = (F(a) - F(b)) / (2 * h);
        const double err = fabs(fd - g[i]), tol = 2e-3 * (fabs(fd) + fabs(g[i])) + 1e-4;
        if (!(err <= tol)) { printf("FAIL %s[%zu]: analytic %.6g fd %.6g\n", name, i, g[i], fd); ++fails; }
    }
}
static float3 v3(const std::vector<double> &x, int o) { return f3((float)x[o], (float)x[o + 1], (float)x[o + 2]); }
static double d3(const std::vector<double> &x, int o, const double *w) { return x[o] * w[0] + x[o + 1] * w[1] + x[o + 2] * w[2]; }

int main() {
    srand(7);
    for (int trial = 0; trial < 20; ++trial) {
        {   // normalize
            std::vector<double> x = {rnd() * 3, rnd() * 3, rnd() * 3};
            const double w[3] = {rnd(), rnd(), rnd()};
            auto F = [&](const std::vector<double> &y) { double n = sqrt(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]); return (y[0] * w[0] + y[1] * w[1] + y[2] * w[2]) / n; };
            float3 g = normalize_vjp(v3(x, 0), f3((float)w[0], (float)w[1], (float)w[2]));
            check("normalize", F, x, {g.x, g.y, g.z});
        }
        {   // ray / triangle
            std::vector<double> x = {rnd(), rnd(), 5 + rnd(), 2 + rnd(), rnd(), rnd(), rnd(), 2 + rnd(), rnd(), 0.3 * rnd(), 0.3 * rnd(), 0, 0.2 * rnd(), 0.2 * rnd(), 1};
            const double gu = rnd(), gv = rnd(), gt = rnd();
            auto F = [&](const std::vector<double> &y) {
                double p0[3] = {y[0], y[1], y[2]}, e1[3] = {y[3], y[4], y[5]}, e2[3] = {y[6], y[7], y[8]}, o[3] = {y[9], y[10], y[11]}, d[3] = {y[12], y[13], y[14]};
                double h[3] = {d[1] * e2[2] - d[2] * e2[1], d[2] * e2[0] - d[0] * e2[2], d[0] * e2[1] - d[1] * e2[0]};
                double a = e1[0] * h[0] + e1[1] * h[1] + e1[2] * h[2], f = 1 / a;
                double s[3] = {o[0] - p0[0], o[1] - p0[1], o[2] - p0[2]};
                double u = f * (s[0] * h[0] + s[1] * h[1] + s[2] * h[2]);
                double q[3] = {s[1] * e1[2] - s[2] * e1[1], s[2] * e1[0] - s[0] * e1[2], s[0] * e1[1] - s[1] * e1[0]};
                double v = f * (d[0] * q[0] + d[1] * q[1] + d[2] * q[2]), t = f * (e2[0] * q[0] + e2[1] * q[1] + e2[2] * q[2]);
                return gu * u + gv * v + gt * t;
            };
            RayTriGrad g = ray_intersect_triangle_vjp(v3(x, 0), v3(x, 3), v3(x, 6), v3(x, 9), v3(x, 12), (float)gu, (float)gv, (float)gt);
            check("ray_tri", F, x, {g.p0.x, g.p0.y, g.p0.z, g.e1.x, g.e1.y, g.e1.z, g.e2.x, g.e2.y, g.e2.z, g.o.x, g.o.y, g.o.z, g.d.x, g.d.y, g.d.z});
        }
        {   // connection factor
            std::vector<double> x = {rnd(), rnd(), rnd(), 3 + rnd(), 2 + rnd(), rnd(), rnd(), 1 + rnd(), rnd(), rnd(), -1 + 0.3 * rnd(), rnd(), 1.0 + 0.1 * rnd()};
            auto F = [&](const std::vector<double> &y) {
                double dv[3] = {y[3] - y[0], y[4] - y[1], y[5] - y[2]};
                double r2 = dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2], r = sqrt(r2);
                double A = (dv[0] * y[6] + dv[1] * y[7] + dv[2] * y[8]) / r, B = (dv[0] * y[9] + dv[1] * y[10] + dv[2] * y[11]) / r;
                return A * fabs(B) / r2 * y[12];
            };
            const double gc = 0.5 + rnd();
            ConnGrad g = connection_vjp(v3(x, 0), v3(x, 3), v3(x, 6), v3(x, 9), (float)x[12], (float)gc);
            std::vector<double> ga = {g.p.x, g.p.y, g.p.z, g.q.x, g.q.y, g.q.z, g.sh_n.x, g.sh_n.y, g.sh_n.z, g.n_q.x, g.n_q.y, g.n_q.z, g.J};
            auto Fs = [&](const std::vector<double> &y) { return gc * F(y); };
            check("connection", Fs, x, ga);
            const float val = connection_value(v3(x, 0), v3(x, 3), v3(x, 6), v3(x, 9), (float)x[12]);
            if (fabs(val - F(x)) > 1e-4 * (1 + fabs(val))) { printf("FAIL connection_value\n"); ++fails; }
        }
        {   // shading normal
            std::vector<double> x = {rnd(), rnd(), 1 + rnd(), rnd(), 1 + rnd(), rnd(), 1 + rnd(), rnd(), rnd(), 0.3 + 0.1 * rnd(), 0.3 + 0.1 * rnd()};
            const double w[3] = {rnd(), rnd(), rnd()};
            auto F = [&](const std::vector<double> &y) {
                double m[3];
                for (int k = 0; k < 3; ++k) m[k] = y[k] + y[9] * (y[3 + k] - y[k]) + y[10] * (y[6 + k] - y[k]);
                double n = sqrt(m[0] * m[0] + m[1] * m[1] + m[2] * m[2]);
                return (m[0] * w[0] + m[1] * w[1] + m[2] * w[2]) / n;
            };
            TriGrad tg; float gu = 0, gv = 0;
            shading_normal_vjp(v3(x, 0), v3(x, 3), v3(x, 6), (float)x[9], (float)x[10], f3((float)w[0], (float)w[1], (float)w[2]), tg, gu, gv);
            check("shading_normal", F, x, {tg.n0.x, tg.n0.y, tg.n0.z, tg.n1.x, tg.n1.y, tg.n1.z, tg.n2.x, tg.n2.y, tg.n2.z, gu, gv});
        }
        {   // face: fn, area, and the vertex-normal share
            std::vector<double> x = {1 + rnd(), rnd(), rnd(), rnd(), 1 + rnd(), rnd()};
            const double wf[3] = {rnd(), rnd(), rnd()}, wc[3] = {rnd(), rnd(), rnd()}, wa = rnd();
            auto F = [&](const std::vector<double> &y) {
                double c[3] = {y[1] * y[5] - y[2] * y[4], y[2] * y[3] - y[0] * y[5], y[0] * y[4] - y[1] * y[3]};
                double len = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
                return d3({c[0] / len, c[1] / len, c[2] / len}, 0, wf) + wa * len * 0.5 + d3({c[0], c[1], c[2]}, 0, wc);
            };
            float3 g1 = f3(0.f), g2 = f3(0.f);
            face_vjp(v3(x, 0), v3(x, 3), f3((float)wf[0], (float)wf[1], (float)wf[2]), (float)wa, f3((float)wc[0], (float)wc[1], (float)wc[2]), g1, g2);
            check("face", F, x, {g1.x, g1.y, g1.z, g2.x, g2.y, g2.z});
        }
    }
    printf(fails ? "adjoint_check: %d FAILURES\n" : "adjoint_check: ok\n", fails);
    return fails ? 1 : 0;
}
