import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import generate_mesh, parse_profile
from bench_ops import meshgen_boundary
from p1afempy_b200 import _lib, cubature
from p1afempy_b200.device import DeviceMesh, context
level = int(sys.argv[1])
c, e = generate_mesh(level)
ctx, dev = context(0); lib = _lib.load()
mesh = DeviceMesh(c, e)
w, p = cubature.resolve_rule(cubature.CubatureRuleEnum.CONICAL_GAUSS_4X4)
xq = mesh.quadrature_points(p); fq = torch.sin(xq[..., 0]) + xq[..., 1]; del xq
u = torch.sin(mesh.coords[:, 0]); fc = torch.ones(mesh.n_elements, dtype=torch.float64, device='cuda')
b = torch.from_numpy(np.vstack(meshgen_boundary(level)).astype(np.int32)).cuda()
none = torch.zeros((0, 2), dtype=torch.int32, device='cuda')
def run(name, fn, reps=3):
    fn(); torch.cuda.synchronize()
    lib.p1_profile_reset(ctx); lib.p1_profile_enable(ctx, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    buf = C.create_string_buffer(1 << 14); lib.p1_profile_report(ctx, buf, len(buf)); lib.p1_profile_enable(ctx, 0)
    print('%-28s %.3f ms' % (name, e0.elapsed_time(e1) / reps), {k: round(v[0] / reps, 3) for k, v in parse_profile(buf.value.decode()).items()}, flush=True)
def lv(): mesh.close(); return mesh.load_vector(fq, w, p)
def gd(): mesh.close(); return mesh.geometric_data([b])
def er(): mesh.close(); return mesh.eta_r(u, b, none, None, fc)
run('load vector cold', lv); run('geometric data cold', gd); run('eta_r cold', er)
fcen = torch.sin(mesh.centroids()[:, 0])
def lm(): mesh.close(); return mesh.load_vector_midpoint(fcen)
run('midpoint load vector cold', lm)
q2, q3, q4 = fq + 1., fq * 2., fq - 1.
def gs(): return mesh.general_stiffness_csr(fq, q2, q3, q4, w, fused=False)
def wm(): return mesh.weighted_mass_csr(fq, w, p)
def it(): return mesh.integrate(fq, w)
def gsf(): return mesh.general_stiffness_csr(fq, q2, q3, q4, w, fused=True)
run('general stiffness cold two-step', gs); run('general stiffness cold fused', gsf); run('weighted mass cold', wm); run('integrate', it)
def ms(): return mesh.mass_csr()
run('mass cold', ms)
rm = DeviceMesh(c, e, reuse=True)
def lvr(): return rm.load_vector(fq, w, p)
def lmr(): return rm.load_vector_midpoint(fcen)
run('load vector, plan reused', lvr); run('midpoint load vector, plan reused', lmr)
