"""The boundary is a C ABI: include/lux_b200.h must compile as plain C99 and a C program must link against
libluxb.so and call the host-only entry points (no GPU needed)."""
import os
import subprocess

import lux_b200 as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

C_SRC = r'''
#include <stdio.h>
#include <string.h>
#include "lux_b200.h"
int main(void) {
  /* the 5-edge graph of tests/golden/hand5.lux.hex: row_end = {2,3,5,5} */
  luxb_eid row_end[4] = {2, 3, 5, 5};
  luxb_vid rl[2], rr[2];
  luxb_eid cl[2];
  int found = luxb_partition_csc(4, 5, row_end, 2, rl, rr, cl);
  printf("%s|found=%d|%u-%u|%u-%u|%llu\n", luxb_version(), found, rl[0], rr[0], rl[1], rr[1], (unsigned long long)cl[1]);
  luxb_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.app = LUXB_PAGERANK; cfg.nranks = 1;
  luxb_csc csc = {4, 5, row_end, NULL, NULL};
  luxb_graph* g = NULL;
  int rc = luxb_open_csc(&csc, &cfg, &g);       /* src == NULL with ne > 0: must be refused, not crash */
  printf("rc=%d err=%s\n", rc, luxb_last_error());
  return (found == 1 && rc < 0) ? 0 : 1;
}
'''


def test_header_is_c99_and_links(tmp_path):
    lib = L.library_path()
    L.load_library()
    src = tmp_path / "abi.c"
    src.write_text(C_SRC)
    exe = tmp_path / "abi"
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src),
                           "-o", str(exe), "-L", os.path.dirname(lib), "-lluxb", "-Wl,-rpath," + os.path.dirname(lib)])
    p = subprocess.run([str(exe)], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "sm_100a" in p.stdout and "found=1|0-2|3-3|5" in p.stdout and "rc=-1" in p.stdout
