"""Run-time options of the drop-in layer.

callbacks = 'device' (default) user callbacks (f, g, a_ij, phi) are first tried
                     on a device-resident ndarray stand-in
                     (p1afempy_b200.devarray): no point array crosses PCIe (at 64M
                     triangles and 16 points that is 17 GB down and 34 GB up for
                     the generalised stiffness).  Arithmetic callbacks (+ - * /,
                     integer powers, sqrt) give the reference's bits;
                     transcendentals come from CUDA's libm and agree with numpy's
                     to ~1e-15 relative instead of bit for bit.  Callbacks the
                     stand-in cannot express fall back to the host one by one.
callbacks = 'host'   the bit-exact opt-in: callbacks are evaluated by numpy on the
                     host on exactly the arrays the reference would pass them;
                     results are bit-identical to the reference wherever the
                     assembly is.
Set with ``p1afempy_b200.options.callbacks = 'host'`` or P1_CALLBACKS=host.

plan_flags           diagnostic P1_PLAN_FORCE_* / P1_PLAN_TINY_SPILL bits (include/
                     p1b200.h) OR-ed into every plan the drop-in layer creates:
                     tests and measurements use them to choose a storage the
                     library would not pick by itself.  Results do not change.
"""
import os

callbacks = os.environ.get('P1_CALLBACKS', 'device')
plan_flags = 0
