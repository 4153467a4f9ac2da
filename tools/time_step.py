"""ms/step of the graph-captured 2 x 128^3 bf16 train step (device-timed, no extras): quick A/B of environment knobs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_segmentation_project_b200.dp import DataParallelTrainer
from multimodal_segmentation_project_b200.models.unet import UNet3D
from multimodal_segmentation_project_b200.synthetic import structured_volume
from multimodal_segmentation_project_b200.utils import metrics as M
dev = torch.device("cuda"); torch.manual_seed(0)
model = UNet3D(1, 4, dropout_rate=0.0).to(dev).train()
tr = DataParallelTrainer(model, M.combined_loss, lr=1e-3, autocast_dtype=torch.bfloat16, metrics_fn="confusion")
x, y = structured_volume(2, 128, seed=1234)
x, y = x.to(dev).bfloat16(), y.to(dev).to(torch.uint8)
tr.capture(x, y, warmup=3)
for _ in range(5): tr.replay()
torch.cuda.synchronize()
K = int(sys.argv[1]) if len(sys.argv) > 1 else 40
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(K): tr.replay()
e.record(); torch.cuda.synchronize()
print(f"{a.elapsed_time(e) / K:.4f} ms/step  env: " + " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("B200_")))
