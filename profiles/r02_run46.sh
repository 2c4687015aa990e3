#!/bin/bash
# round 2, call 50: the launch-bound case (shape_rope at the reference's size): differentiated scan eager vs two CUDA graphs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 170 python profiles/micro/graph_scan_timing.py 2> gpurun_out/r02_46_graph.err | tail -1 > gpurun_out/r02_46_graph.json
cat gpurun_out/r02_46_graph.json; tail -3 gpurun_out/r02_46_graph.err
