mkdir -p gpurun_out
rm -f gpurun_out/exp_stem2.jsonl gpurun_out/exp_stem2.err
run() { env "$@" timeout 120 python scripts/layer_times.py 512 3 "$*" >> gpurun_out/exp_stem2.jsonl 2>> gpurun_out/exp_stem2.err; }
run OGL_FUSE_STEM=1
run OGL_FUSE_STEM=1 OGL_DBG=64
run OGL_FUSE_STEM=1 OGL_DBG=128
run OGL_FUSE_STEM=1 OGL_DBG=1
run OGL_FUSE_STEM=1 OGL_DBG=4
run OGL_FUSE_STEM=1 OGL_DBG=5
run OGL_FUSE_STEM=1 OGL_DBG=69
python scripts/show_exp.py gpurun_out/exp_stem2.jsonl | cut -c1-60,200-; tail -3 gpurun_out/exp_stem2.err
