"""Run a PySPH-style Application script on the B200 path, unmodified:

    python -m rigid_body_2d_3d_pysph_b200.run code/benchmark_2_....py --tf 0.1
"""
import os
import runpy
import sys

from .compat.install import install


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        print(__doc__)
        return 2
    script, rest = argv[0], argv[1:]
    install()
    # as `python script.py` would: the script's siblings (its own geometry.py)
    # import by bare name
    sys.path.insert(0, os.path.dirname(os.path.abspath(script)))
    sys.argv = [script] + rest
    runpy.run_path(script, run_name='__main__')
    return 0


if __name__ == '__main__':
    sys.exit(main())
