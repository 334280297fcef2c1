# Convenience targets; the driver uses __graft_entry__.build() / pytest / bench.py directly.
PY ?= python

.PHONY: build oracle test test-gpu bench bench-ref clean

build:            ## nvcc (sm_100a) -> monte_carlo_retirement_b200/_lib/libmcr_b200.so, gcc -> oracle
	$(PY) -c "import __graft_entry__ as g; g.build()"

oracle:
	$(MAKE) -C oracle

test:             ## CPU suite (oracle pins, host logic, ABI, gloo sharding, payload / report, fuzz pins)
	$(PY) -m pytest tests -x -q -m "not gpu"

test-gpu:         ## parity suite, needs a B200
	$(PY) -m pytest tests -x -q -m gpu

bench:
	$(PY) bench.py --gpus 1 --steps 10 --warmup 3

bench-ref:        ## the CPU arm (oracle port on all host cores)
	$(PY) bench.py --impl reference --steps 2 --warmup 1

clean:
	rm -rf monte_carlo_retirement_b200/_lib oracle/_build
