PY ?= python

.PHONY: build test test-gpu bench sass clean lint docs doctor

build:
	$(PY) -c "import __graft_entry__ as g; g.build()"

test: build
	$(PY) -m pytest tests/ -x -q -m "not gpu"

test-gpu: build
	$(PY) -m pytest tests/ -x -q -m gpu

bench: build
	$(PY) bench.py

sass: build
	cuobjdump -sass bagua_b200/_C.so > /tmp/bagua_b200.sass && grep -c "UTCHMMA\|UTMALDG" /tmp/bagua_b200.sass

docs:
	$(PY) scripts/gen_api_docs.py > docs/api.md

doctor: build
	$(PY) -m bagua_b200.script.bagua_doctor

lint:
	$(PY) -m pyflakes bagua_b200 tests bench.py || true

clean:
	rm -rf bagua_b200/csrc/build bagua_b200/_C.so bagua_b200/_C_torch.so bagua_b200/libnccl-net-bagua.so bagua_b200/*.stamp
