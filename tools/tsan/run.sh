#!/bin/sh
# ThreadSanitizer pass over the host-only threaded code: the table writer, the bfh parser and the
# expansion pool (the pool's class is cut out of scasa_cuda.cu, which cannot be built without nvcc).
# Usage: sh tools/tsan/run.sh      (prints "ok" / "pool ok"; any data race is reported by TSan)
set -e
here=$(cd "$(dirname "$0")" && pwd); root=$(cd "$here/../.." && pwd)
out=/tmp/scasa_tsan; mkdir -p $out
python3 - "$root" "$out" <<'PY'
import sys
root, out = sys.argv[1], sys.argv[2]
s = open(root + "/scasa_b200/csrc/scasa_cuda.cu").read()
a = s.index("class ExpandPool {"); b = s.index("\n};\n", a) + 4
open(out + "/pool_snippet.h", "w").write(s[a:b])
PY
src=$root/scasa_b200/csrc
g++ -std=c++17 -O1 -g -fsanitize=thread -I$root/include -I$src $here/host_driver.cpp $src/scasa_write.cpp $src/scasa_bfh.cpp $src/scasa_host.cpp -lpthread -o $out/host_driver
g++ -std=c++17 -O1 -g -fsanitize=thread -march=native -I$root/include -I$src -I$out $here/pool_driver.cpp $src/scasa_host.cpp -lpthread -o $out/pool_driver
$out/host_driver
$out/pool_driver
# AddressSanitizer + UBSan: the bfh / bustools readers and the writers through their Python tests,
# against a host-only build of those sources
g++ -std=c++17 -O1 -g -fPIC -shared -fsanitize=address,undefined -I$root/include -I$src $here/host_stubs.cpp $src/scasa_write.cpp $src/scasa_bfh.cpp $src/scasa_bus.cpp $src/scasa_host.cpp -lpthread -o $out/libhost_asan.so
LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) ASAN_OPTIONS=detect_leaks=0 python3 $here/run_asan_tests.py 2>&1 | grep -v "Warning\|sparse_coo"
