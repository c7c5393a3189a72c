#!/bin/bash
# The GPU sessions that produced profiles/*_r02* (run each part under gpurun; keep gpurun_out/ below 64 MiB per call):
#   a: ncu launch list + ncu --set full captures      -> tools/final_session_a.sh
#   d: parity tests, re-capture of the top kernel, bench line + reference arm -> tools/final_session_d.sh
#   e: whole GPU suite + bench line on the final code -> tools/final_session_e.sh
# then here: python tools/make_profiles.py r02
set -u
bash tools/final_session_a.sh && bash tools/final_session_d.sh && bash tools/final_session_e.sh
