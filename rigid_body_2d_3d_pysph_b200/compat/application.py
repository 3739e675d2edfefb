"""``pysph.solver.application.Application`` surface.

[upstream, restated] hook order of ``Application.run``:
initialize -> create_scheme (in __init__) ; parse argv (application +
scheme options, conflict_handler='resolve') ; consume_user_options ;
scheme.consume_user_options ; configure_scheme ; create_particles ;
create_equations ; solver.setup ; solve (post_step hook every step).
"""
import argparse
import os
import sys

from . import output as _output
from .scheme import Scheme


class Application(object):
    def __init__(self, fname=None, output_dir=None, domain=None):
        self.domain = domain
        self.solver = None
        self.particles = []
        self.options = None
        self.args = None
        self.scheme = None
        if fname is None:
            main = sys.modules.get('__main__')
            f = getattr(main, '__file__', None) or 'output'
            fname = os.path.splitext(os.path.basename(f))[0]
        self.fname = fname
        self.output_dir = os.path.abspath(output_dir or fname + '_output')
        self.initialize()
        self.scheme = self.create_scheme()
        self._setup_argparse()

    # -- hooks scripts override ------------------------------------------------
    def initialize(self):
        pass

    def add_user_options(self, group):
        pass

    def consume_user_options(self):
        pass

    def create_scheme(self):
        return None

    def configure_scheme(self):
        pass

    def create_particles(self):
        raise NotImplementedError()

    def create_equations(self):
        return self.scheme.get_equations()

    def create_solver(self):
        return self.scheme.get_solver()

    def create_domain(self):
        return None

    def create_nnps(self):
        return None

    def create_inlet_outlet(self, particle_arrays):
        return []

    def create_tools(self):
        return []

    def pre_step(self, solver):
        pass

    def post_stage(self, current_time, dt, stage):
        pass

    def post_step(self, solver):
        pass

    def post_process(self, info_fname_or_directory):
        pass

    def customize_output(self):
        pass

    def _mayavi_config(self, code):
        pass

    # ----------------------------------------------------------------------------
    def _setup_argparse(self):
        p = argparse.ArgumentParser(
            description=self.__doc__, conflict_handler='resolve')
        p.add_argument('--openmp', action='store_true', dest='with_openmp',
                       default=None, help='accepted for compatibility; the '
                       'path runs on the GPU')
        p.add_argument('--no-openmp', action='store_false',
                       dest='with_openmp')
        p.add_argument('--opencl', action='store_true', default=False)
        p.add_argument('--cuda', action='store_true', default=False)
        p.add_argument('--tf', '--final-time', dest='final_time', type=float,
                       default=None)
        p.add_argument('--timestep', '--time-step', dest='time_step',
                       type=float, default=None)
        p.add_argument('--max-steps', dest='max_steps', type=int, default=None)
        p.add_argument('--pfreq', dest='freq', type=int, default=None)
        p.add_argument('-d', '--directory', dest='output_dir',
                       default=self.output_dir)
        p.add_argument('--fname', dest='fname', default=self.fname)
        p.add_argument('--disable-output', dest='disable_output',
                       action='store_true', default=False)
        p.add_argument('--detailed-output', dest='detailed_output',
                       action='store_true', default=False)
        p.add_argument('--no-graph', dest='use_graph', action='store_false',
                       default=True, help='do not capture steps in CUDA '
                       'graphs')
        p.add_argument('-q', '--quiet', action='store_true', default=False)
        user = p.add_argument_group('User', 'User defined command line '
                                    'arguments')
        self.add_user_options(user)
        if self.scheme is not None:
            sg = p.add_argument_group('Scheme options')
            self.scheme.add_user_options(sg)
        self.arg_parse = p

    def _parse(self, argv):
        if argv is None:
            argv = sys.argv[1:]
        self.options = self.arg_parse.parse_args(argv)
        self.args = argv

    def setup(self, argv=None):
        self._parse(argv)
        o = self.options
        self.output_dir = os.path.abspath(o.output_dir)
        self.fname = o.fname
        self.consume_user_options()
        if self.scheme is not None:
            self.scheme.consume_user_options(o)
            self.configure_scheme()
        self.particles = self.create_particles()
        self.solver = self.create_solver()
        s = self.solver
        if o.final_time is not None:
            s.set_final_time(o.final_time)
        if o.time_step is not None:
            s.set_time_step(o.time_step)
        if o.freq is not None:
            s.set_print_freq(o.freq)
        if o.max_steps is not None:
            s.set_max_steps(o.max_steps)
        s.set_disable_output(o.disable_output)
        s.detailed_output = o.detailed_output
        s.use_graph = o.use_graph
        s.set_output_directory(self.output_dir)
        s.set_output_fname(self.fname)
        if type(self).post_step is not Application.post_step:
            s.add_post_step_callback(self.post_step)
        if type(self).pre_step is not Application.pre_step:
            s.add_pre_step_callback(self.pre_step)
        self.equations = self.create_equations()
        s.setup(self.particles, self.equations, kernel=s.kernel)

    def run(self, argv=None):
        self.setup(argv)
        self.solver.solve(not self.options.quiet)
        self.info_filename = os.path.join(self.output_dir,
                                          self.fname + '.info')
        os.makedirs(self.output_dir, exist_ok=True)
        with open(self.info_filename, 'w') as f:
            f.write('{"completed": true, "fname": "%s", "output_dir": "%s"}\n'
                    % (self.fname, self.output_dir))

    @property
    def output_files(self):
        if self.solver is not None and self.solver.output_files:
            return list(self.solver.output_files)
        if os.path.isdir(self.output_dir):
            return _output.get_files(self.output_dir, self.fname)
        return []
