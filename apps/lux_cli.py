"""CLI drivers with the reference's flags and stdout lines:

  python apps/lux_cli.py pagerank   -ng 1 -ni 10 -file g.lux [-verbose]            # pagerank/pagerank.cc:121-148
  python apps/lux_cli.py components -ng 2 -file g.lux [-check] [-verbose]          # components/components.cc:145-175
  python apps/lux_cli.py sssp       -ng 1 -file g.lux -start 0 [-check]            # sssp/sssp.cc
  python apps/lux_cli.py colfilter  -ng 1 -ni 10 -file ratings.lux                 # col_filter/colfilter.cc:85-107
  python apps/lux_cli.py converter  -nv N -ne M -input edges.txt -output g.lux     # tools/converter.cc:13-39 (host only)

`-ll:gpu N` is accepted as a synonym of `-ng N` (README.md:47); -ll:fsize / -ll:zsize are accepted and ignored (HBM is
managed by the library).  With -ng > 1 the driver re-launches itself under torch.distributed.run, one rank per GPU.
Prints the reference's lines: "[Memory Setting] Set ll:fsize >= %zuMB and ll:zsize >= %zuMB" (pagerank.cc:84-85,
components.cc:87-88), "ELAPSED TIME = %7.7f s" (pagerank.cc:118), "[PASS]/[FAIL] Check task: rowLeft(%u)
numMistakes(%u)" (components_gpu.cu:831-836).  `-out file.npy` additionally saves the vertex values (the reference never writes its results anywhere).
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

APPS = {"pagerank": 0, "components": 1, "sssp": 2, "colfilter": 3}


def parse(argv):
    opt = dict(ng=1, ni=10, file=None, start=0, verbose=False, check=False, out=None)
    i = 0
    while i < len(argv):
        a = argv[i]
        if a in ("-ng", "-ll:gpu"):
            opt["ng"] = int(argv[i + 1]); i += 1
        elif a == "-ni":
            opt["ni"] = int(argv[i + 1]); i += 1
        elif a == "-file":
            opt["file"] = argv[i + 1]; i += 1
        elif a == "-start":
            opt["start"] = int(argv[i + 1]); i += 1
        elif a == "-out":
            opt["out"] = argv[i + 1]; i += 1
        elif a in ("-verbose", "-v"):
            opt["verbose"] = True
        elif a in ("-check", "-c"):
            opt["check"] = True
        elif a in ("-ll:fsize", "-ll:zsize", "-ll:cpu", "-ll:util"):
            i += 1  # Legion/Realm memory flags: accepted, not needed
        i += 1
    return opt


def memory_setting(app, nv, ne, bounds, frontier_bytes):
    """The reference's advice formulas, verbatim in sizes: pagerank.cc:61-85, components.cc:57-88."""
    V, E, VTX = 4, 8, (80 if app == "colfilter" else 4)
    max_fb, max_edges = 0, 0
    P = len(bounds["row_left"])
    for p in range(P):
        nodes = int(bounds["row_right"][p]) - int(bounds["row_left"][p]) + 1
        nodes = max(nodes, 0) if nodes < (1 << 31) else 0
        nxt = int(bounds["col_left"][p + 1]) if p + 1 < P else ne
        edges = max(nxt - int(bounds["col_left"][p]), 0)
        if app in ("pagerank", "colfilter"):
            edge_struct, node_struct = (12 if app == "colfilter" else 8), 16
            fb = edges * edge_struct + nodes * node_struct + nodes * V + nodes * VTX + nv * VTX
        else:
            fb = edges * 8 + edges * 4 + nodes * 8 + nv * 8 + nodes * 2 * VTX + nv * VTX + frontier_bytes * 2
        max_fb, max_edges = max(max_fb, fb), max(max_edges, edges)
    if app in ("pagerank", "colfilter"):
        zc = ne * V + nv * E + nv * V + nv * 2 * VTX
    else:
        zc = ne * V + nv * E + nv * 2 * VTX + frontier_bytes * 2 + nv * 8 + max_edges * 4
    return max_fb // 1024 // 1024 + 1, zc // 1024 // 1024 + 1


def converter(argv):
    """tools/converter.cc: same flags, same first stdout line; the conversion itself is luxb_convert_edgelist."""
    nv, ne, inp, out = 0, 0, "", ""
    i = 0
    while i < len(argv):
        if argv[i] == "-nv":
            nv = int(argv[i + 1]); i += 1
        elif argv[i] == "-ne":
            ne = int(argv[i + 1]); i += 1
        elif argv[i] == "-input":
            inp = argv[i + 1]; i += 1
        elif argv[i] == "-output":
            out = argv[i + 1]; i += 1
        i += 1
    print("nv = %d ne = %d input = %s output = %s" % (nv, ne, inp, out), flush=True)  # converter.cc:80
    import lux_b200 as L
    try:
        L.convert_edgelist(inp, out, nv, ne)
    except L.LuxError as e:
        print("converter: %s" % e, file=sys.stderr)
        return 1
    return 0


def main():
    if len(sys.argv) >= 2 and sys.argv[1] == "converter":
        return converter(sys.argv[2:])
    if len(sys.argv) < 2 or sys.argv[1] not in APPS:
        print(__doc__)
        return 2
    app = sys.argv[1]
    opt = parse(sys.argv[2:])
    if not opt["file"]:
        print("Missing -file (graph in .lux format, see tools/converter.cc)")
        return 2
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if opt["ng"] > 1 and world == 1:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(opt["ng"]),
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("LUX_PORT", "29611")] + sys.argv
        return subprocess.call(cmd)
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    import lux_b200 as L
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g = L.LuxGraph.from_file(opt["file"], app=APPS[app], rank=rank, nranks=world, device=local, start=opt["start"],
                             verbose=opt["verbose"])
    b = g.bounds()
    if rank == 0:
        fb, zc = memory_setting(app, g.nv, g.ne, b, int(b["fq_right"][-1]) + 1)
        print("[Memory Setting] Set ll:fsize >= %dMB and ll:zsize >= %dMB" % (fb, zc), flush=True)
    g.comm_init_torch()
    g.init()
    if app in ("pagerank", "colfilter"):
        g.iterate(opt["ni"])
    else:
        g.run_to_convergence()
    if rank == 0:
        print("ELAPSED TIME = %7.7f s" % g.stats()["loop_seconds"], flush=True)
    if opt["check"] and app in ("components", "sssp"):
        bad = g.check()
        print("[%s] Check task: rowLeft(%u) numMistakes(%u)" % ("PASS" if bad == 0 else "FAIL", int(b["row_left"][rank]), bad),
              flush=True)
    if opt["out"]:
        vals = g.values()
        if rank == 0:
            np.save(opt["out"], vals)
    g.close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
