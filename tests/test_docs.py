"""Every ```python block of docs/guide/*.md is executed: the blocks of one chapter share a namespace and run in order,
so the guide cannot drift from the code."""
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHAPTERS = sorted(glob.glob(os.path.join(ROOT, "docs", "guide", "*.md")))
BLOCK = re.compile(r"```python\n(.*?)```", re.S)


@pytest.mark.parametrize("path", CHAPTERS, ids=[os.path.basename(p) for p in CHAPTERS])
def test_guide_chapter_runs(path, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv("DGL_DOWNLOAD_DIR", str(tmp_path / "dgl_home"))
    monkeypatch.setenv("DGL_DIST_MODE", "standalone")          # chapters may set it; restored after the test
    text = open(path).read()
    blocks = BLOCK.findall(text)
    assert blocks, f"{path}: no python blocks"
    ns = {"__name__": "__guide__"}
    for i, code in enumerate(blocks):
        try:
            exec(compile(code, f"{os.path.basename(path)}[block {i}]", "exec"), ns)
        except Exception as e:  # noqa: BLE001
            raise AssertionError(f"{os.path.basename(path)} block {i} failed: {type(e).__name__}: {e}\n{code}") from e


def test_guide_is_indexed():
    index = open(os.path.join(ROOT, "docs", "index.md")).read()
    for path in CHAPTERS:
        assert os.path.basename(path) in index, f"{os.path.basename(path)} is not linked from docs/index.md"
