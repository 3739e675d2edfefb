"""``pysph.sph.scheme`` surface: Scheme, SchemeChooser, add_bool_argument.

[upstream, restated] SURVEY.md App. C-8.  ``_smart_getattr`` is the piece that
makes the command-line defaults win over constructor arguments (divergence D6:
effective kf = 1e3, fric_coeff = 0.5 whatever the script passes;
/root/reference/code/rigid_body_3d.py:633-636).
"""


def add_bool_argument(group, arg, dest, help, default=True):
    group.add_argument('--%s' % arg, action='store_true', dest=dest,
                       help=help)
    neg_help = 'Do not ' + help[0].lower() + help[1:]
    group.add_argument('--no-%s' % arg, action='store_false', dest=dest,
                       help=neg_help)
    group.set_defaults(**{dest: default})


class Scheme(object):
    def __init__(self, fluids=None, solids=None, dim=2):
        self.fluids = fluids
        self.solids = solids
        self.dim = dim
        self.solver = None

    def add_user_options(self, group):
        pass

    def attributes_changed(self):
        pass

    def configure(self, **kw):
        for k, v in kw.items():
            if not hasattr(self, k):
                raise RuntimeError('Parameter %s not defined for %s.' %
                                   (k, self.__class__.__name__))
            setattr(self, k, v)
        self.attributes_changed()

    def consume_user_options(self, options):
        pass

    def configure_solver(self, kernel=None, integrator_cls=None,
                         extra_steppers=None, **kw):
        raise NotImplementedError()

    def get_equations(self):
        raise NotImplementedError()

    def get_solver(self):
        return self.solver

    def setup_properties(self, particles, clean=True):
        pass

    def _smart_getattr(self, obj, var):
        res = getattr(obj, var, None)
        if res is None:
            return getattr(self, var)
        return res


class SchemeChooser(Scheme):
    def __init__(self, default, **schemes):
        self.default = default
        self.schemes = dict(schemes)
        self.scheme = schemes[default]

    def add_user_options(self, group):
        for scheme in self.schemes.values():
            try:
                scheme.add_user_options(group)
            except Exception as e:  # argparse conflict: two schemes, same flag
                if 'conflicting option' not in str(e):
                    raise
        choices = list(self.schemes.keys())
        group.add_argument('--scheme', action='store', dest='scheme',
                           default=self.default, choices=choices,
                           help='Specify scheme to use (one of %s).' %
                           choices)

    def attributes_changed(self):
        self.scheme.attributes_changed()

    def configure(self, **kw):
        self.scheme.configure(**kw)

    def consume_user_options(self, options):
        self.scheme = self.schemes[options.scheme]
        self.scheme.consume_user_options(options)

    def configure_solver(self, kernel=None, integrator_cls=None,
                         extra_steppers=None, **kw):
        self.scheme.configure_solver(kernel=kernel,
                                     integrator_cls=integrator_cls,
                                     extra_steppers=extra_steppers, **kw)

    def get_equations(self):
        return self.scheme.get_equations()

    def get_solver(self):
        return self.scheme.get_solver()

    def setup_properties(self, particles, clean=True):
        self.scheme.setup_properties(particles, clean)
