import sys, os, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
from qkan_implementation_b200.degree_optimizer import ChebyshevLeastSquares
n, F, D = 774_456, 79, 3
gen = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((n, F), dtype=torch.float64, device="cuda", generator=gen) * 0.6
y = (torch.cos(2 * x[:, 0]) + 0.3 * x[:, 1] ** 3 + 0.1 * torch.randn(n, dtype=torch.float64, device="cuda", generator=gen)).contiguous()
w = torch.rand(n, dtype=torch.float64, device="cuda", generator=gen) + 0.5
eng = ChebyshevLeastSquares(D)
eng.solve(x, y, w)
coef, ybar = eng.last["coef"], eng.last["ybar"]
def timeit(f, reps=5):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
print("gram (+ D2H)            %.2f ms" % timeit(lambda: eng.gram(x, y).cpu()))
print("residuals with X^T r    %.2f ms" % timeit(lambda: eng.residual_sums(x, y, w, coef, ybar, True)))
print("residuals, sums only    %.2f ms" % timeit(lambda: eng.residual_sums(x, y, w, coef, ybar, False)))
G = eng.last["gram"]; P = F * (D + 1)
t0 = time.perf_counter()
for _ in range(5):
    fs = [eng._pinv_factor(G[:F * (d + 1), :F * (d + 1)], n) for d in range(D + 1)]
print("host eigen-factors      %.2f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
print("solve total             %.2f ms" % timeit(lambda: eng.solve(x, y, w)))
