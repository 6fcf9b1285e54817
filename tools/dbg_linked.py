"""debug: random linked model seed S on the device vs the oracle (heap_left, inbox counts)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import happysim_b200 as hs
from happysim_b200 import _abi as A
from happysim_b200.linked import LinkedRun
import oracle_lib as O, random_models as RM
for seed in [int(x) for x in sys.argv[1:]] or [0]:
    lm, end_s, what = RM.random_linked_model(seed)
    end_ns, nP, n = int(end_s * 1e9), lm.n_partitions, 6
    caps = [dict(record_cap=512, sample_cap=64, service_cap=64) for _ in range(nP)]
    run = LinkedRun(lm)
    outs, (delivered, lost, over) = run.run(seed=1000 + seed, end_ns=end_ns, n_replicas=n, caps=caps)
    inb = [e.read_box("inbox")[1] if lm.models[q].inbox_cap else np.zeros(n, np.uint32) for q, e in enumerate(run.engines)]
    run.close()
    ps = [O.make_params(seed=1000 + seed, end_ns=end_ns, n_replicas=n, rid_base=q, rid_stride=nP + 1, **caps[q]) for q in range(nP)]
    want, wd, wl, _ = O.oracle_run_linked(lm, ps, end_ns=end_ns, cseed=1000 + seed)
    print(os.path.basename(os.environ.get("HS_B200_LIB", "default")), what)
    print(" delivered dev", delivered.tolist(), "oracle", wd.tolist(), "lost", lost.tolist(), wl.tolist(), "over", over.tolist())
    print(" caps", [(m.outbox_cap, m.inbox_cap) for m in lm.models], "windows", run.windows)
    for q in range(nP):
        print(" part", q, "dev heap_left", outs[q]["summaries"]["heap_left"].tolist(), "inbox", inb[q].tolist(),
              "oracle", want[q]["summaries"]["heap_left"].tolist(),
              "status", outs[q]["summaries"]["status"].tolist(), want[q]["summaries"]["status"].tolist(), "ev", outs[q]["summaries"]["events_processed"].tolist()[:3], want[q]["summaries"]["events_processed"].tolist()[:3], "ev equal", bool((outs[q]["summaries"]["events_processed"] == want[q]["summaries"]["events_processed"]).all()),
              "final_time dev", outs[q]["summaries"]["final_time_ns"].tolist()[:2], "or", want[q]["summaries"]["final_time_ns"].tolist()[:2])
