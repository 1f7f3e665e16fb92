"""CPU checks of the CLI front-end: flag parsing keeps the reference's spellings, and the [Memory Setting] advice
reproduces the reference's formulas (pagerank.cc:61-85, components.cc:57-88) on a hand-computable case."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("lux_cli", os.path.join(ROOT, "apps", "lux_cli.py"))
cli = importlib.util.module_from_spec(spec)
spec.loader.exec_module(cli)


def test_flag_spellings():
    o = cli.parse(["-ll:gpu", "4", "-ll:fsize", "12000", "-ll:zsize", "20000", "-ni", "7", "-file", "g.lux", "-v", "-c", "-start", "9"])
    assert o["ng"] == 4 and o["ni"] == 7 and o["file"] == "g.lux" and o["verbose"] and o["check"] and o["start"] == 9
    o = cli.parse(["-ng", "2", "-file", "x.lux", "-verbose", "-check"])
    assert o["ng"] == 2 and o["verbose"] and o["check"] and o["ni"] == 10


def test_memory_setting_formulas():
    MB = 1024 * 1024
    nv, ne = 1000 * MB // 4, 2000 * MB // 4  # sizes chosen so every term is a whole number of MB
    b = dict(row_left=np.array([0], np.uint32), row_right=np.array([nv - 1], np.uint32), col_left=np.array([0], np.uint64))
    # pagerank.cc:77-85: fb = ne*8 + nv*16 + nv*4 + nv*4 + nv*4 ; zc = ne*4 + nv*8 + nv*4 + nv*2*4
    fb, zc = cli.memory_setting("pagerank", nv, ne, b, 0)
    assert fb == (ne * 8 + nv * 28) // MB + 1 and zc == (ne * 4 + nv * 20) // MB + 1
    # components.cc:70-85 with frontierSize F
    F = 8 + 4 * ((nv - 1) // 16 + 100)
    fb, zc = cli.memory_setting("components", nv, ne, b, F)
    assert fb == (ne * 8 + ne * 4 + nv * 8 + nv * 8 + nv * 8 + nv * 4 + 2 * F) // MB + 1
    assert zc == (ne * 4 + nv * 8 + nv * 8 + 2 * F + nv * 8 + ne * 4) // MB + 1


def test_converter_tool_runs_without_a_gpu(tmp_path):
    """`lux_cli.py converter` = tools/converter.cc (flags -nv -ne -input -output, first stdout line) on the product's
    host-only luxb_convert_edgelist; the file it writes loads through the oracle's reader."""
    import subprocess
    import sys
    import oracle as O
    edges = [(0, 1), (1, 2), (2, 0), (3, 0), (0, 2)]
    txt = tmp_path / "e.txt"
    txt.write_text("".join("%d %d\n" % e for e in edges))
    out = str(tmp_path / "g.lux")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "apps", "lux_cli.py"), "converter", "-nv", "4", "-ne", "5", "-input", str(txt),
                        "-output", out], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert p.stdout.startswith("nv = 4 ne = 5 input = %s output = %s" % (txt, out))
    row_end, src = O.lux_read(out)
    assert row_end.tolist() == [2, 3, 5, 5] and src.tolist() == [2, 3, 0, 0, 1]
