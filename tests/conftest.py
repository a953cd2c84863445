import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _assets():
    staged = os.path.join(ROOT, "assets", "_ref")
    if os.path.isdir(os.path.join(staged, "onnx")):
        return staged
    if os.path.isdir("/root/reference/3rdparty/onnx"):
        return "/root/reference/3rdparty"
    return None


@pytest.fixture(scope="session")
def assets_dir():
    d = _assets()
    if d is None:
        pytest.skip("model/image assets not staged (run __graft_entry__.build() where /root/reference exists)")
    return d


@pytest.fixture(scope="session")
def sad_linus_full(assets_dir):
    from zaru_b200.synth import load_image_rgba
    return load_image_rgba(os.path.join(assets_dir, "img", "sad_linus.jpg"))


@pytest.fixture(scope="session")
def sad_linus_cropped(assets_dir):
    from zaru_b200.synth import load_image_rgba
    return load_image_rgba(os.path.join(assets_dir, "img", "sad_linus_cropped.jpg"))
