"""Where config 5's encode time goes: H2D alone, encoder alone (eager / graph), per tower."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cmh_b200.valid_loop import DchmtModel
from cmh_b200.codes import CodeBuffer
dev = torch.device("cuda", 0)
B, ctx = int(os.environ.get("B", 500)), 32
torch.manual_seed(1)
model = DchmtModel(64).to(dev).to(torch.bfloat16).eval()
img_h = torch.randn(B, 3, 224, 224).to(torch.bfloat16).pin_memory()
txt_h = torch.randint(1, 49000, (B, ctx)); txt_h[:, -1] = 49407; txt_h = txt_h.pin_memory()
img_d, txt_d = img_h.to(dev), txt_h.to(dev)
def t(fn, n=10):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
dst = torch.empty_like(img_d)
print("H2D image batch ms", round(t(lambda: dst.copy_(img_h, non_blocking=True)), 3), "GB/s", round(img_h.numel() * 2 / 1e6 / t(lambda: dst.copy_(img_h, non_blocking=True)), 1))
with torch.no_grad():
    print("image tower eager ms", round(t(lambda: model.clip.encode_image(img_d)), 3))
    print("text tower eager ms", round(t(lambda: model.clip.encode_text(txt_d)), 3))
    buf = CodeBuffer(B, 64, dev); idx = torch.arange(B, device=dev)
    w, b = model.image_hash.weight.float().contiguous(), model.image_hash.bias.float().contiguous()
    f = model.image_hash.hidden(model.clip.encode_image(img_d))
    print("head kernel ms", round(t(lambda: buf.put_head(idx, f, w, b, relu=True)), 4))
    s = torch.cuda.Stream(dev); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        model.clip.encode_image(img_d); model.clip.encode_text(txt_d)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        o1 = model.clip.encode_image(img_d); o2 = model.clip.encode_text(txt_d)
        buf.put_head(idx, model.image_hash.hidden(o1), w, b, relu=True)
    print("both towers + head, graph replay ms", round(t(lambda: g.replay()), 3))
    flops = B * (4.37e9 + 2.9e9)
    print("=> TFLOP/s at graph rate", round(flops / (t(lambda: g.replay()) * 1e-3) / 1e12, 1))
