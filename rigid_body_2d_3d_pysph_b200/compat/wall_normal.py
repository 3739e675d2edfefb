"""``pysph.sph.isph.wall_normal`` descriptors (ComputeNormals, SmoothNormals).

[upstream, restated] the arithmetic is the in-tree EDAC copy at
/root/reference/code/boundary_particles.py:71-135 with the upstream property
names ``normal_tmp`` / ``normal``; it is evaluated by
compat.sph_evaluator.SPHEvaluator at setup time.
"""
from .equation import Equation


class ComputeNormals(Equation):
    pass


class SmoothNormals(Equation):
    pass
