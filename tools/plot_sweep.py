#!/usr/bin/env python
"""Plot of a bench_sweep JSON (the counterpart of the reference's results/random_qp_benchmark plot,
benchmarks/random_qps.py:83-97): solve time against problem size on log-log axes for the GPU kernels (fp64, fp32)
and the CPU oracle, written as a dependency-free SVG.

    python tools/plot_sweep.py profiles/r02_sweep.json profiles/r02_sweep.svg"""
import json
import math
import sys

src, dst = sys.argv[1], sys.argv[2]
d = json.load(open(src))
series = {}
for r in d["single"]:
    series.setdefault("B200 " + r["dtype"], []).append((r["nx"], r["solve_ms"]))
    if "cpu_us_per_iter" in r:
        series.setdefault("CPU oracle " + r["dtype"], []).append((r["nx"], r["cpu_us_per_iter"] * r["cpu_iters"] * 1e-3))
W, Hh, ml, mb, mt, mr = 720, 460, 70, 50, 30, 170
xs = [p[0] for s in series.values() for p in s]
ys = [p[1] for s in series.values() for p in s]
x0, x1 = math.log10(min(xs)) - 0.05, math.log10(max(xs)) + 0.05
y0, y1 = math.floor(math.log10(min(ys))), math.ceil(math.log10(max(ys)))
px = lambda x: ml + (math.log10(x) - x0) / (x1 - x0) * (W - ml - mr)
py = lambda y: Hh - mb - (math.log10(y) - y0) / (y1 - y0) * (Hh - mb - mt)
col = {"B200 f64": "#1f77b4", "B200 f32": "#2ca02c", "CPU oracle f64": "#d62728", "CPU oracle f32": "#ff7f0e"}
o = ['<svg xmlns="http://www.w3.org/2000/svg" width="%d" height="%d" font-family="sans-serif" font-size="12">' % (W, Hh),
     '<rect width="100%" height="100%" fill="white"/>',
     '<text x="%d" y="18" font-size="14">Solve time, rand_qp(nx, nx/4, nx/4), eps_abs 1e-3, cold start (%s)</text>' % (ml, d.get("gpu", ""))]
for e in range(y0, y1 + 1):
    y = py(10.0 ** e)
    o.append('<line x1="%d" y1="%.1f" x2="%d" y2="%.1f" stroke="#ddd"/><text x="%d" y="%.1f" text-anchor="end">1e%d ms</text>' % (ml, y, W - mr, y, ml - 6, y + 4, e))
for x in sorted(set(xs)):
    o.append('<line x1="%.1f" y1="%d" x2="%.1f" y2="%d" stroke="#eee"/><text x="%.1f" y="%d" text-anchor="middle">%d</text>' % (px(x), mt, px(x), Hh - mb, px(x), Hh - mb + 16, x))
o.append('<text x="%d" y="%d" text-anchor="middle">nx (D = 2 nx)</text>' % ((ml + W - mr) // 2, Hh - 12))
for i, (name, pts) in enumerate(sorted(series.items())):
    pts = sorted(pts)
    c = col.get(name, "#555")
    o.append('<polyline fill="none" stroke="%s" stroke-width="2" points="%s"/>' % (c, " ".join("%.1f,%.1f" % (px(x), py(y)) for x, y in pts)))
    for x, y in pts:
        o.append('<circle cx="%.1f" cy="%.1f" r="3" fill="%s"/>' % (px(x), py(y), c))
    o.append('<rect x="%d" y="%d" width="14" height="4" fill="%s"/><text x="%d" y="%d">%s</text>' % (W - mr + 12, mt + 20 * i + 8, c, W - mr + 32, mt + 20 * i + 14, name))
o.append("</svg>")
open(dst, "w").write("\n".join(o))
print("wrote", dst)
