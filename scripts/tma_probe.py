"""Feed rate of a TMA + mbarrier ring with nothing consuming the data (v4h_debug_tma_probe): bytes per cycle
and SM as a function of box height, boxes per stage, issuing warps, ring depth, number of CTAs and working
set."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit4hep_b200 import _cabi
lib = _cabi.load(); dev = torch.device("cuda:0")
s = torch.cuda.current_stream().cuda_stream
clk_ghz = 1.965

def run(buf, rows, cols, stages, boxes, box_rows, producers, ctas):
    iters = max(50, 9600 * 1024 // (boxes * box_rows * 128))
    cyc = torch.zeros(ctas, dtype=torch.int64, device=dev)
    for _ in range(2):
        _cabi.check(lib.v4h_debug_tma_probe(buf.data_ptr(), rows, cols, stages, boxes, box_rows, producers, iters, ctas, cyc.data_ptr(), s))
    torch.cuda.synchronize()
    return iters * boxes * box_rows * 128 / cyc.float().mean().item()

# row pitches that are NOT powers of two (a 2 KB / 8 KB pitch camps on L2 slices and caps the rate at 36 B/clk/SM)
for ws_mb, rows, cols in ((16, 8192, 960), (528, 65536, 4032), (16, 8192, 1024)):
    buf = torch.zeros(rows, cols, dtype=torch.bfloat16, device=dev)
    for ctas in (148,):
        # (box_rows, boxes per stage): 48 KB stages cut in different ways, then 16 KB stages
        for box_rows, boxes in ((256, 1), (128, 3), (64, 6), (128, 1), (64, 2)):
            for producers in (1, 2, 4):
                if producers > boxes: continue
                stages = min(8, (200 * 1024) // (boxes * box_rows * 128))
                bpc = run(buf, rows, cols, stages, boxes, box_rows, producers, ctas)
                print(f"ws {ws_mb:4d} MB, {ctas} CTAs, stage = {boxes} x [{box_rows} x 64] ({boxes * box_rows // 8} KB), S={stages}, "
                      f"{producers} issuing warp(s): {bpc:5.1f} B/clk/SM = {bpc * ctas * clk_ghz / 1e3:5.2f} TB/s", flush=True)
