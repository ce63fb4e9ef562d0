"""Device-side transform_x alone at the bench batch (64 raw 1024x1024 RGB images -> 336): CUDA-event timing and the
HBM fraction on its algorithmic bytes (3*H0*W0 read + 12*S*S written per image)."""
import sys, json, torch
sys.path.insert(0, ".")
from aaclip_b200 import ops
B, H0, W0, S = 64, int(sys.argv[1]) if len(sys.argv) > 1 else 1024, int(sys.argv[2]) if len(sys.argv) > 2 else 1024, 336
x = torch.randint(0, 256, (B, H0, W0, 3), dtype=torch.uint8, device="cuda")
for _ in range(3):
    y = ops.preprocess_u8(x, S)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
e0.record()
for _ in range(n):
    y = ops.preprocess_u8(x, S)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
byts = B * (3 * H0 * W0 + 12 * S * S)
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
print(f"transform_x {B} x {H0}x{W0} -> {S}: {ms * 1e3:.1f} us, {B / ms * 1e3:.0f} img/s, {byts / ms / 1e6:.0f} GB/s algorithmic "
      f"({byts / ms / 1e6 / peak:.3f} of {peak} GB/s)")
