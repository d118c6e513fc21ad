#!/bin/bash
# SASS mnemonic counts of the hot kernels -> profiles/r02_sass_summary.txt (evidence of TMA / bulk-copy instructions)
cuobjdump -sass fen_b200/libfen_gpu.so 2>/dev/null | python scripts/sass_count.py > /tmp/sass.txt
(echo "# SASS mnemonic counts per kernel of fen_b200/libfen_gpu.so (cuobjdump -sass, sm_100a), round 2."
 echo "# UTMALDG = TMA tensor load (cp.async.bulk.tensor), UBLKCP = bulk copy shared -> global (cp.async.bulk, the slab transposes' peer stores),"
 echo "# SYNCS = mbarrier ops, LDGSTS = cp.async, BAR.SYNC = block / named barriers.  Generated on the build box: scripts/sass_summary.sh"
 cat /tmp/sass.txt) > profiles/r02_sass_summary.txt
