#!/bin/sh
# Differential test of the number formatter (scasa_format_real15) against the version at a git revision:
#   sh tools/fmt_diff/run.sh [rev = HEAD~1]
# 19 M values (gamma-distributed estimates, all decimal exponents, random bit patterns, rounded values, neighbours of
# powers of ten); any difference in the text is printed.  Used when the formatter's arithmetic is reorganised
# (the powl table and llrintl: 0 differences).
set -e
rev=${1:-HEAD~1}
here=$(cd "$(dirname "$0")" && pwd); root=$(cd "$here/../.." && pwd)
out=/tmp/scasa_fmt_diff; mkdir -p $out
src=$root/scasa_b200/csrc
git -C $root show $rev:scasa_b200/csrc/scasa_write.cpp | sed 's/scasa_format_real15/old_format/g' > $out/old_w.cpp
sed 's/scasa_format_real15/new_format/g' $src/scasa_write.cpp > $out/new_w.cpp
for v in old new; do
    g++ -O2 -std=c++17 -fPIC -shared -I$root/include -I$src $out/${v}_w.cpp $root/tools/tsan/host_stubs.cpp $src/scasa_host.cpp -o $out/lib$v.so -lpthread
done
g++ -O2 -std=c++17 -w $here/cmp.cpp -L$out -lold -lnew -Wl,-rpath,$out -o $out/cmp
$out/cmp
