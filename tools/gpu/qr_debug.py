import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rust-la_b200", "python"))
import numpy as np
np.set_printoptions(linewidth=200, precision=6)
from oracle import oracle as o
from la import Matrix, QRDecomposition
o.build()
for a in [np.array([[12.0, -51.0, 4.0], [6.0, 167.0, -68.0], [-4.0, 24.0, -41.0]]), o.fill((5, 5), 1), o.fill((200, 200), 1), o.fill((300, 200), 1)]:
    m, n = a.shape
    qr = QRDecomposition.new(Matrix.from_numpy(a))
    p, rd = o.qr(a)
    got = qr.get_qr().to_numpy()
    print("shape", a.shape, "packed err", np.abs(got - p).max(), "rdiag err", np.abs(qr.rdiag - rd).max())
    q, r = qr.get_q().to_numpy(), qr.get_r().to_numpy()
    rq = o.qr_get_q(p, rd)
    print("  q err", np.abs(q - rq).max(), "r err", np.abs(r - o.qr_get_r(p, rd)).max(), "qr-a", np.abs(q @ r - a).max())
    if m <= 5:
        print(got, qr.rdiag, "\nq=", q, "\nrefq=", rq, "\nr=", r)
        t = np.empty(128 * 128)
        qr._tmat_buf.download(t)
        print("Tt block:", t.reshape(128, 128)[:m, :m])
