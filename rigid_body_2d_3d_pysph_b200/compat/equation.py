"""``pysph.sph.equation`` surface: Equation / Group / MultiStageEquations.

In PySPH these carry per-particle Python that is transpiled to Cython.  Here
they are *descriptors*: the product's equations
(rigid_body_2d_3d_pysph_b200.rigid_body_common) keep the reference's class
names and constructor arguments, and the integrator maps the group list that
``Scheme.get_equations()`` returns onto CUDA launches (see
rigid_body_2d_3d_pysph_b200.compat.integrator.plan_from_equations).  An
equation the planner does not know is an error -- there is no interpreted
fallback.  (Group semantics: SURVEY.md App. C-4.)
"""


class Equation(object):
    def __init__(self, dest, sources):
        self.dest = dest
        self.sources = sources if sources is not None and len(sources) > 0 \
            else None
        self.name = self.__class__.__name__
        self.var_name = ''

    def __repr__(self):
        return '%s(dest=%r, sources=%r)' % (self.name, self.dest,
                                            self.sources)


class Group(object):
    def __init__(self, equations, real=True, update_nnps=False, iterate=False,
                 max_iterations=1, min_iterations=0, pre=None, post=None,
                 condition=None, start_idx=0, stop_idx=None, name=None):
        self.equations = equations
        self.real = real
        self.update_nnps = update_nnps
        self.iterate = iterate
        self.max_iterations = max_iterations
        self.min_iterations = min_iterations
        self.pre = pre
        self.post = post
        self.condition = condition
        self.name = name

    def __repr__(self):
        return 'Group(%r, real=%r)' % (self.equations, self.real)


class MultiStageEquations(object):
    def __init__(self, groups):
        self.groups = groups

    def __len__(self):
        return len(self.groups)

    def __getitem__(self, i):
        return self.groups[i]
