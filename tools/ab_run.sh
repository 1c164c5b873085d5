#!/bin/bash
# On the GPU box: time the AMP kernel of every experiment library given (tags; "tag:ENV=VAL" sets an env var).
for spec in "$@"; do
  t=${spec%%:*}; envs=""
  if [[ "$spec" == *:* ]]; then envs=${spec#*:}; fi
  echo "== $spec"
  env ${envs//,/ } SPARC_B200_LIB=build/lib_$t.so python tools/profile_amp.py --T 8 --launches 4 2>&1 | tail -1
done
