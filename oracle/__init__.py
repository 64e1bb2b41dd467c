"""CPU oracle — TEST INFRASTRUCTURE ONLY.

Restates, on CPU, what the reference (aligokalppeker/hypelcnn) computes on the hot path.
Nothing under hypelcnn_b200/ imports this package; only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs do.

Parity status: index arithmetic, graph structure, scene preparation, patch gather and the
confusion-matrix metrics are PINNED bit-exactly to outputs of the reference's own code
(tests/golden/make_golden.py).  The floating-point semantics inside TensorFlow ops are a
restatement of the published behaviour of tensorflow 2.9 / tf-slim 1.1.0 and are
"parity unpinned" (TensorFlow cannot be installed here; the reference has no tests).
"""
