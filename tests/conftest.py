import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def write_plot_stubs(directory):
    """Minimal stand-ins for the plotting packages qgs.diagnostics imports at module level (matplotlib, IPython,
    ipywidgets are not in the image); nothing of them is used by the computations under test."""
    import os
    files = {"matplotlib/__init__.py": "", "matplotlib/pyplot.py": "def get_cmap(name=None):\n    return name\n",
             "matplotlib/animation.py": "", "matplotlib/ticker.py": "def FuncFormatter(f):\n    return f\n",
             "IPython/__init__.py": "", "IPython/display.py": "HTML = display = None\n",
             "ipywidgets/__init__.py": "interactive = None\n"}
    for rel, text in files.items():
        path = os.path.join(str(directory), rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as fh:
            fh.write(text)
    return str(directory)
