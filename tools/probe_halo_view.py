"""Probe behind csrc/conv_halo.cu: single-tap identity filters show which pixel / channel group each output reads when the
nine taps are shifted shared-memory descriptor views of one TMA-loaded halo tile (run on a B200: python tools/probe_halo_view.py)."""
import sys; sys.path.insert(0, '.')
import numpy as np, torch
from cnn_lstm_ctc_ocr_b200 import _lib as L
lib = L.load()
B, H, W, C, Co = 1, 16, 8, 32, 32
def conv(x, w, path):
    L.check(lib.ocr_conv_set_path(path), "p")
    dx = torch.tensor(x, device="cuda"); dw = torch.tensor(np.ascontiguousarray(w.reshape(9 * C, Co).T), device="cuda")
    out = torch.zeros((B, H, W, Co), device="cuda"); zb = torch.zeros(Co, device="cuda")
    L.check(lib.ocr_conv3x3_same(L.ptr(dx), B, H, W, C, L.ptr(dw), L.ptr(zb), Co, 0, L.ptr(out), L.stream_handle()), "c")
    torch.cuda.synchronize(); L.check(lib.ocr_conv_set_path(0), "p")
    return out.cpu().numpy()
yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
pix = (yy * 16 + xx + 1).astype(np.float32)                       # pixel id (0 = padding)
for tap in (4, 0, 1, 3, 5, 8):
    w = np.zeros((3, 3, C, Co), np.float32)
    w[tap // 3, tap % 3] = np.eye(C, Co)
    xp = np.broadcast_to(pix[None, :, :, None], (B, H, W, C)).astype(np.float32).copy()
    o = conv(xp, w, 2)[0, :, :, 0]
    ref = conv(xp, w, 1)[0, :, :, 0]
    print("tap", tap, "pixel-id view (halo kernel), rows 0..3:"); print(o[:4].astype(int)); print("gather kernel:"); print(ref[:4].astype(int))
    xc = np.broadcast_to(np.arange(C, dtype=np.float32)[None, None, None, :], (B, H, W, C)).copy()
    oc = conv(xc, w, 2)[0, 2, :, :]      # row 2: [x, co] -> which channel arrived
    print("  channel view, y=2, x=0..7 (first 12 co):"); print(oc[:, :12].astype(int))
