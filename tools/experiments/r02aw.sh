#!/bin/bash
set -x
O=gpurun_out
python -m pytest tests -x -q -m gpu > $O/aw_pytest.log 2>&1; echo "rc=$?" >> $O/aw_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/aw_smoke.log 2>&1; echo "rc=$?" >> $O/aw_smoke.log
