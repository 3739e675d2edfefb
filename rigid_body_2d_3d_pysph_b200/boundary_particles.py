"""Boundary-particle identification -- same names as the reference's
``code/boundary_particles.py``.  Setup-time only: it produces ``is_boundary``,
which the scripts copy into ``contact_force_is_boundary``, the source gate of
the contact loop (rigid_body_common.py:678).  Evaluated on the host by
compat.sph_evaluator (vectorised NumPy over cKDTree neighbour lists).
"""
from .compat.equation import Equation, Group
from .compat.wall_normal import ComputeNormals, SmoothNormals


def add_boundary_identification_properties(pa):
    """boundary_particles.py:9-19"""
    pa.add_property('normal', stride=3)
    pa.add_property('normal0', stride=3)
    pa.add_property('normal_tmp', stride=3)
    pa.add_property('normal_norm')
    pa.add_property('is_boundary', type='int')
    pa.add_output_arrays(['is_boundary'])


class IdentifyBoundaryParticleCosAngle(Equation):
    """boundary_particles.py:22-68"""

    def __init__(self, dest, sources):
        super(IdentifyBoundaryParticleCosAngle, self).__init__(dest, sources)


def get_boundary_identification_etvf_equations(destinations, sources,
                                               boundaries=None):
    """boundary_particles.py:190-216"""
    eqs = []
    g1, g2, g3 = [], [], []
    all = list(set(destinations + sources))
    for dest in destinations:
        g1.append(ComputeNormals(dest=dest, sources=all))
    for dest in destinations:
        g2.append(SmoothNormals(dest=dest, sources=[dest]))
    for dest in destinations:
        if boundaries is None:
            srcs = [dest]
        else:
            srcs = list(set([dest] + boundaries))
        g3.append(IdentifyBoundaryParticleCosAngle(dest=dest, sources=srcs))
    eqs.append(Group(equations=g1))
    eqs.append(Group(equations=g2))
    eqs.append(Group(equations=g3))
    return eqs
