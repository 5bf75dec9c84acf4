#!/bin/sh
# oracle/build_reference_drivers.sh -- TEST INFRASTRUCTURE (called by oracle/Makefile, target _ref/qphandler_hs071).
# Builds the programs in which the reference's OWN translation units drive the CUDA plugins:
#   _ref/qphandler_hs071[_twin]                 src/QPhandler.cpp on the plugins (qphandler_test.cpp)
#   _ref/algorithm_nl[_twin|_twin_noclip]       src/Algorithm.cpp + src/SQPTNLP.cpp + src/QPhandler.cpp on the plugins (algorithm_test.cpp)
# The reference sources are copied to a scratch directory under /tmp, patched there with integration/restartsqp_cuda_backend.patch
# and compiled from there: no reference source enters the repository.  qpOASES / QORE / Ipopt are absent: stubs_link holds
# declaration-level stand-ins whose solver entry points abort.  `_twin`: linked with the CPU twin of the C ABI (capi_twin.cpp over
# liboracle.so) instead of libsqpb200.so.  `_noclip`: -DSQPB200_QORE_NO_CLIP (see tests/test_reference_algorithm.py).
set -e
REF=${1:-/root/reference}
CXX=${CXX:-g++}
HERE=$(cd "$(dirname "$0")" && pwd)
ADAPTER=$HERE/../restartsqp_b200/csrc/adapter
W=$(mktemp -d /tmp/sqpb200_patched_reference.XXXXXX)
trap 'rm -rf "$W"' EXIT
cp -r "$REF/include" "$REF/src" "$W/" && chmod -R u+w "$W"
(cd "$W" && patch -p1 -s -i "$HERE/../integration/restartsqp_cuda_backend.patch")
cp "$ADAPTER/CudaQPInterface.hpp" "$ADAPTER/CudaQOREInterface.hpp" "$W/include/sqphot/"
mkdir -p "$W/o" "$W/n" "$HERE/_ref"
INC="-I$HERE/stubs_link -I$HERE/stubs -I$W/include -I$W/include/sqphot -I$HERE/../include -I$HERE"
FLAGS="-O2 -std=c++11 -w"
# translation units that see CudaQOREInterface.hpp are compiled twice (with and without the clipping of its setters)
SEES="$W/src/QPhandler.cpp $W/src/Algorithm.cpp $ADAPTER/CudaQOREInterface.cpp $HERE/algorithm_test.cpp $HERE/qphandler_test.cpp"
REST="$W/src/SQPTNLP.cpp $W/src/qpOASESInterface.cpp $W/src/QOREInterface.cpp $W/src/Options.cpp $W/src/Utils.cpp $W/src/Vector.cpp $W/src/SpTripletMat.cpp $W/src/SpHbMat.cpp $ADAPTER/CudaQPInterface.cpp $HERE/stubs_link/link_standins.cpp $HERE/capi_twin.cpp"
{
  for f in $SEES $REST; do echo "$CXX $FLAGS $INC -c $f -o $W/o/$(basename $f .cpp).o"; done
  for f in $SEES; do echo "$CXX $FLAGS -DSQPB200_QORE_NO_CLIP $INC -c $f -o $W/n/$(basename $f .cpp).o"; done
} | xargs -P 8 -I CMD sh -c CMD
O=$W/o
COMMON="$O/QPhandler.o $O/qpOASESInterface.o $O/QOREInterface.o $O/Options.o $O/Utils.o $O/Vector.o $O/SpTripletMat.o $O/SpHbMat.o $O/CudaQPInterface.o $O/CudaQOREInterface.o $O/link_standins.o"
PRODUCT="-L$HERE/../restartsqp_b200/lib -lsqpb200 -ldl -Wl,-rpath,\$ORIGIN/../../restartsqp_b200/lib"
TWIN="$O/capi_twin.o -L$HERE -loracle -ldl -Wl,-rpath,\$ORIGIN/.."
$CXX -o "$HERE/_ref/qphandler_hs071" $O/qphandler_test.o $COMMON $PRODUCT
$CXX -o "$HERE/_ref/qphandler_hs071_twin" $O/qphandler_test.o $COMMON $TWIN
$CXX -o "$HERE/_ref/algorithm_nl" $O/algorithm_test.o $O/Algorithm.o $O/SQPTNLP.o $COMMON $PRODUCT
$CXX -o "$HERE/_ref/algorithm_nl_twin" $O/algorithm_test.o $O/Algorithm.o $O/SQPTNLP.o $COMMON $TWIN
N=$W/n
$CXX -o "$HERE/_ref/algorithm_nl_twin_noclip" $N/algorithm_test.o $N/Algorithm.o $O/SQPTNLP.o $N/QPhandler.o $O/qpOASESInterface.o $O/QOREInterface.o $O/Options.o \
    $O/Utils.o $O/Vector.o $O/SpTripletMat.o $O/SpHbMat.o $O/CudaQPInterface.o $N/CudaQOREInterface.o $O/link_standins.o $TWIN
# the reference's real qpOASESInterface.cpp on the functional qpOASES stand-in (oracle solver) beside the plugin on the CPU twin
$CXX $FLAGS -DQPOASES_OVER_ORACLE $INC -c "$HERE/stubs_link/link_standins.cpp" -o "$W/o/link_standins_f.o"
$CXX $FLAGS $INC -c "$HERE/stubs_link/qpoases_over_oracle.cpp" -o "$W/o/qpoases_over_oracle.o"
$CXX $FLAGS $INC -I"$ADAPTER" -c "$HERE/backend_pin_test.cpp" -o "$W/o/backend_pin_test.o"
$CXX -o "$HERE/_ref/backend_pin_twin" $O/backend_pin_test.o $O/qpOASESInterface.o $O/qpoases_over_oracle.o $O/link_standins_f.o $O/Options.o $O/Utils.o $O/Vector.o \
    $O/SpTripletMat.o $O/SpHbMat.o $O/CudaQPInterface.o $TWIN
