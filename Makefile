# Convenience targets (everything is also reachable through __graft_entry__.build()).
PY ?= python

build:
	$(MAKE) -C pd_mg_pin_corrosion_b200/csrc
	$(MAKE) -C host
	$(MAKE) -C oracle port
	@if [ -d /root/reference/src ]; then $(MAKE) -C oracle ref; fi

test-cpu:
	$(PY) -m pytest tests -q -m "not gpu"

test-gpu:
	$(PY) -m pytest tests -q -m gpu

bench:
	$(PY) bench.py

.PHONY: build test-cpu test-gpu bench
