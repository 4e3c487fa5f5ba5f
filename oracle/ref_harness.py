"""Import the UNMODIFIED reference (/root/reference) in the build container.

TEST INFRASTRUCTURE ONLY (see oracle/rt_oracle.c header).  The reference's
environment.py imports gymnasium, stable_baselines3 and matplotlib (via
visualize_voxel.py), none of which are installed here and none of which take part
in the arithmetic of the hot path, so they are replaced by inert stub modules.
draw_line.py and transforms.py import as they are.  The reference reads ./data
relative to the CWD (environment.py:28-29,94), so the import happens with the CWD
switched to the reference root and the class keeps absolute paths afterwards.

This module only works where /root/reference exists (the build container); it is
used by oracle/gen_golden.py to produce tests/golden/*.npz and by the optional
`reference`-marked tests.  Nothing here travels to the GPU box at run time.
"""
import contextlib
import os
import sys
import types

REF_ROOT = os.environ.get("RT_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "environment.py"))


def _install_stubs() -> None:
    if "gymnasium" not in sys.modules:
        gym = types.ModuleType("gymnasium")

        class Env:  # gymnasium.Env: no behaviour used by the reference beyond the base
            def __init__(self, *a, **k):
                pass

        class Box:  # gymnasium.spaces.Box: shape/dtype holder
            def __init__(self, low, high, shape=None, dtype=None):
                self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

        spaces = types.ModuleType("gymnasium.spaces")
        spaces.Box = Box
        gym.Env = Env
        gym.spaces = spaces
        sys.modules["gymnasium"] = gym
        sys.modules["gymnasium.spaces"] = spaces
    if "stable_baselines3" not in sys.modules:
        sb3 = types.ModuleType("stable_baselines3")
        common = types.ModuleType("stable_baselines3.common")
        checker = types.ModuleType("stable_baselines3.common.env_checker")
        checker.check_env = lambda env, *a, **k: None
        sb3.common = common
        common.env_checker = checker
        sys.modules["stable_baselines3"] = sb3
        sys.modules["stable_baselines3.common"] = common
        sys.modules["stable_baselines3.common.env_checker"] = checker
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        mpl = types.ModuleType("matplotlib")
        pyplot = types.ModuleType("matplotlib.pyplot")
        widgets = types.ModuleType("matplotlib.widgets")
        widgets.Slider = object
        mpl.pyplot = pyplot
        mpl.widgets = widgets
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = pyplot
        sys.modules["matplotlib.widgets"] = widgets
    if "torchsummary" not in sys.modules:
        ts = types.ModuleType("torchsummary")
        ts.summary = lambda *a, **k: None
        sys.modules["torchsummary"] = ts


@contextlib.contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


_cache = {}


def load():
    """Return a namespace with the reference modules: draw_line, transforms, environment."""
    if _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    _install_stubs()
    sys.path.insert(0, REF_ROOT)
    try:
        with _cwd(REF_ROOT):
            import draw_line
            import transforms
            import environment
    finally:
        sys.path.remove(REF_ROOT)
    ns = types.SimpleNamespace(draw_line=draw_line, transforms=transforms, environment=environment)
    _cache["ns"] = ns
    return ns


def load_networks():
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    _install_stubs()
    sys.path.insert(0, REF_ROOT)
    try:
        import networks
    finally:
        sys.path.remove(REF_ROOT)
    return networks


class RefEnv:
    """The reference RadiotherapyEnv with an explicit tumour choice.

    environment.py:90 draws the tumour file from the global NumPy RNG over an
    unsorted os.listdir; for reproducible traces the draw is replaced by the
    caller's file name (np.random.choice is patched for the duration of reset()).
    Everything else is the reference's own code.
    """

    def __init__(self, visionless=True, tumour_name=None):
        ns = load()
        self._np_random = ns.environment.np.random
        self._next = tumour_name
        with _cwd(REF_ROOT), self._patched_choice():
            self.env = ns.environment.RadiotherapyEnv(visionless=visionless)

    @contextlib.contextmanager
    def _patched_choice(self):
        orig = self._np_random.choice
        if self._next is not None:
            name = self._next
            self._np_random.choice = lambda seq, *a, **k: name
        try:
            yield
        finally:
            self._np_random.choice = orig

    def reset(self, tumour_name=None):
        if tumour_name is not None:
            self._next = tumour_name
        with _cwd(REF_ROOT), self._patched_choice():
            return self.env.reset()

    def step(self, action):
        return self.env.step(action)
