set -e
python tests/perf_sweep_vs_reference.py --shape reddit --ks 8,16,32,64
python tests/perf_sweep_vs_reference.py --shape flickr --ks 32
python tests/perf_sweep_vs_reference.py --shape yelp --ks 32
python tests/perf_sweep_vs_reference.py --shape proteins --ks 64
python tests/perf_sweep_vs_reference.py --shape products --ks 32
python tests/perf_sweep_vs_reference.py --shape reddit --ks 32 --kind powerlaw
