"""Where the fused transformer stack (csrc/token_tc.cu) spends its time: launch duration against the number of
panoramas (resident groups of 16 CTAs decide the wave count) and clock stamps of one worker thread per GEMM phase."""
import sys

import torch

sys.path.insert(0, ".")
from omnifusion_b200 import _lib
from omnifusion_b200.checkpoint import synthetic_state_dict
from omnifusion_b200.model.spherical_model_iterative import spherical_fusion

DEV = torch.device("cuda:0")
L = _lib.lib()
n_tok = int(sys.argv[1]) if len(sys.argv) > 1 else 18
nrows = {18: 4, 26: 5, 46: 6}[n_tok]
net = spherical_fusion(nrows, n_tok, (128, 128), (80, 80))
net.load_state_dict(synthetic_state_dict("iterative", n_tok, 0))
net = net.to(DEV).eval()
net._ensure_handle(DEV)
net._ensure_weights()
print("resident groups:", L.ofb_token_stack_resident_groups(n_tok))


def launch(B, x, scratch, enc, nblk=6, stop=0):
    _lib.check(L.ofb_token_stack_f32(net._handle, _lib.ptr(x), _lib.ptr(scratch), scratch.numel(), _lib.ptr(enc), B, n_tok,
                                     nblk, stop, _lib.stream_of(DEV)))


for B in (1, 4, 8, 9, 10, 16, 32):
    x0 = torch.randn(B * n_tok, 512, device=DEV)
    scratch = torch.empty(L.ofb_token_stack_scratch_floats(B, n_tok), device=DEV)
    enc = torch.empty(B * n_tok, 512, device=DEV)
    x = x0.clone()
    for _ in range(3):
        launch(B, x, scratch, enc)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        launch(B, x, scratch, enc)
    e1.record()
    torch.cuda.synchronize()
    print(f"B={B:2d}: {e0.elapsed_time(e1) / 20 * 1e3:8.1f} us per launch (6 blocks)")

B = 4
stamps = torch.zeros(24 * 8, dtype=torch.int64, device=DEV)
_lib.check(L.ofb_debug_token_stamps(_lib.ptr(stamps)))
x = torch.randn(B * n_tok, 512, device=DEV)
scratch = torch.empty(L.ofb_token_stack_scratch_floats(B, n_tok), device=DEV)
enc = torch.empty(B * n_tok, 512, device=DEV)
launch(B, x, scratch, enc)
launch(B, x, scratch, enc)
torch.cuda.synchronize()
_lib.check(L.ofb_debug_token_stamps(None))
st = stamps.cpu().reshape(24, 8)
names = {0: ["fill operand (LN1)", "MMAs (qkv)", "epilogue + partial scores", "exchange A", "softmax + PV", "exchange B"],
         1: ["fill operand (att)", "MMAs (proj)", "epilogue", "exchange C"],
         2: ["fill operand (LN2)", "MMAs (fc1)", "epilogue (GELU -> operand)"],
         3: ["-", "MMAs (fc2)", "epilogue (partial sums)", "exchange D", "reduce", "exchange E"]}
for g in range(4, 8):        # second block: steady state
    row = st[g]
    d = [int(row[k + 1] - row[k]) for k in range(len(names[g & 3]))]
    print(f"phase {g & 3}: " + " | ".join(f"{n} {v}" for n, v in zip(names[g & 3], d)))
print("phase 0 detail: partial-score reads + sum -> shared", int(st[4][7] - st[4][4]), "| softmax + PV + store", int(st[4][5] - st[4][7]))
print("block 1 total clk:", int(st[8][0] - st[4][0]), " whole stack:", int(st[23][6] - st[0][0]))
