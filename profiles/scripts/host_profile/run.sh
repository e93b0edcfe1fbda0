#!/bin/bash
# builds the host mirror against the stub engine and times the host side of one sharded generation
set -e
here=$(cd "$(dirname "$0")" && pwd); root=$here/../../..
mkdir -p /tmp/gsb_host_profile
gcc -O2 -g -fPIC -shared -w -o /tmp/gsb_host_profile/libgsb200.so $root/tests/c/stub_engine.c
gcc ${OPT:--O2} -g -std=gnu11 -DGSB_HOST_PROFILE -I$root/include -o /tmp/gsb_host_profile/drive $here/drive.c $root/genomicsimulation_b200/host/gsb_simdata.c \
    -L/tmp/gsb_host_profile -lgsb200 -lm -Wl,-rpath,/tmp/gsb_host_profile
/tmp/gsb_host_profile/drive "$@"
