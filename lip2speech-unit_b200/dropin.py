"""Run an UNMODIFIED reference script with the B200 generator in place of its own class.

    python /path/to/repo/lip2speech-unit_b200/dropin.py multi_input_vocoder/inference_server.py <its arguments ...>
    python /path/to/repo/lip2speech-unit_b200/dropin.py multi_input_vocoder/inference.py <its arguments ...>

The reference scripts do ``from models_multi_input import MelCodeGenerator`` (inference.py:28,
inference_server.py:28) from their own directory; Python puts a script's directory at ``sys.path[0]``,
ahead of ``PYTHONPATH``, so path shadowing alone can never replace a sibling module.  This launcher
therefore registers this package's ``models_multi_input`` in ``sys.modules`` under that name *before*
the script runs (the import system looks there first), then executes the script with ``runpy`` exactly
as ``python script.py`` would: ``__name__ == '__main__'``, ``sys.argv[0]`` = the script, its directory
at ``sys.path[0]``.  Workers forked by ``multiprocessing.Pool`` (inference.py:249) inherit the registration.

``install()`` does the registration only, for callers that import the reference modules themselves.
"""
import importlib.util
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG_NAME = "lip2speech_unit_b200"


def load_package():
    """Import this hyphen-named directory as the package ``lip2speech_unit_b200``."""
    if PKG_NAME in sys.modules:
        return sys.modules[PKG_NAME]
    spec = importlib.util.spec_from_file_location(PKG_NAME, os.path.join(HERE, "__init__.py"),
                                                  submodule_search_locations=[HERE])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[PKG_NAME] = mod
    spec.loader.exec_module(mod)
    return mod


def install():
    """Make ``import models_multi_input`` resolve to the B200 classes, whatever sys.path says."""
    pkg = load_package()
    sys.modules["models_multi_input"] = pkg.models_multi_input
    return pkg.models_multi_input


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ("-h", "--help"):
        sys.stderr.write(__doc__)
        return 2
    script = os.path.abspath(argv[0])
    if not os.path.isfile(script):
        sys.stderr.write(f"dropin: no such script: {argv[0]}\n")
        return 2
    install()
    sys.argv = [script] + argv[1:]
    if sys.path and os.path.abspath(sys.path[0] or os.getcwd()) == HERE:
        sys.path.pop(0)                              # this launcher's own directory, put there by `python dropin.py`
    sys.path.insert(0, os.path.dirname(script))      # what `python script.py` does
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
