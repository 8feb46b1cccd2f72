# Convenience targets (everything is plain python underneath)
PY ?= python

.PHONY: build test test-gpu samples bench flags api sass stress clean

build:            ## compile the sm_100a extension in-tree (cross-compiles without a GPU)
	$(PY) -c "import __graft_entry__ as g; g.build()"

test:             ## CPU suite
	$(PY) -m pytest tests -q -m "not gpu"

test-gpu:         ## on a B200 box
	$(PY) -m pytest tests -q -m gpu

samples:          ## sample inputs for the EXAMPLE USAGE lines (bin/sample_*, bin/inputs/*)
	bin/make_samples.sh

bench:            ## headline benchmark on 1 GPU
	$(PY) bench.py --gpus 1 --steps 20 --warmup 5

flags:            ## regenerate docs/FLAGS.md
	$(PY) -m harmony_b200 flags --markdown > docs/FLAGS.md

api:              ## regenerate docs/API.md
	$(PY) tools/gen_api_doc.py

sass:             ## per-kernel SASS mnemonic summary (offline)
	$(PY) tools/sass_summary.py

stress:           ## concurrent-jobs stress with hang watchdog (host backend)
	$(PY) tools/stress_multijob.py SchedulerImpl 5

clean:
	rm -rf harmony_b200/ops/_build build *.egg-info
