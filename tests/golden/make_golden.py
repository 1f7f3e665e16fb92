"""Regenerates the committed golden fixtures.  Run from the repo root:  python tests/golden/make_golden.py
hand5.lux.hex is NOT produced by the oracle: it is the byte image that the reference's tools/converter.cc writes for
edges {0->1,1->2,2->0,3->0,0->2}: python oracle/build_ref.py && printf '0 1\\n1 2\\n2 0\\n3 0\\n0 2\\n' > e5.txt &&
oracle/_ref/converter -nv 4 -ne 5 -input e5.txt -output e5.lux  (hex dump of e5.lux)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle as O  # noqa: E402

row_end, src = O.gen_rmat_csc(10, 1000, 16000, 27)
ss = O.label_run(O.APP_SSSP, row_end, src, start=0)
np.savez_compressed(os.path.join(HERE, "oracle_rmat10.npz"), row_end=row_end, src=src,
                    pagerank10=O.pagerank(row_end, src, 10), cc=O.label_run(O.APP_CC, row_end, src)["labels"],
                    sssp0=ss["labels"], sssp0_active=ss["active"])
print("wrote oracle_rmat10.npz")
