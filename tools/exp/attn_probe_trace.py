"""Isolated timeline of attn_decode_kernel (no neighbours): umv_op_attention_block at the 14B head geometry, B=8, ctx 1058."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from unimedvl_b200 import config as ucfg, _lib
from unimedvl_b200.engine import Engine
H, HKV, DH = 28, 4, 128
QN = (H + 2 * HKV) * DH
B, ctx = int(os.environ.get("B", 8)), int(os.environ.get("CTX", 1058))
dims = ucfg.BagelDims(llm=ucfg.LLMDims(hidden=H * DH, heads=H, kv_heads=HKV, inter=128, layers=1, vocab=1024), vit=ucfg.ViTDims(hidden=144, heads=2, inter=328, layers=1))
e = Engine(dims, max_tokens=2048, max_seqs=16, kv_pages=int(os.environ.get("PAGES", 512)), enable_vit=False, enable_gen=False)
e.fill_synthetic(1); e.finalize()
seqs = [e.seq_new() for _ in range(B)]
for s in seqs:
    e.attention_block(0, [s], [ctx], list(range(ctx)), qkv=torch.randn(ctx, QN).bfloat16(), update_kv=True)
# flush L2 so the K/V come from HBM as in the decode step
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
part = torch.randn(3, B, QN, device="cuda"); bias = torch.randn(QN).bfloat16().cuda()
for rep in range(3):
    flush.zero_(); torch.cuda.synchronize()
    _lib.check(e.lib.umv_trace_begin(64))
    out, path = e.attention_block(0, seqs, [1] * B, [5] * B, partial=part, bias=bias, update_kv=False)
    st = np.zeros((64, 12), dtype=np.uint64); names = C.create_string_buffer(64 * 32); n = C.c_int32()
    _lib.check(e.lib.umv_trace_read(st.ctypes.data_as(C.c_void_p), names, 32, 64, C.byref(n)))
    _lib.check(e.lib.umv_trace_begin(0))
    t = st[:n.value].astype(np.int64)
    for i in range(n.value):
        nm = names.raw[i * 32:(i + 1) * 32].split(b"\0")[0].decode()
        if nm == "attn_decode":
            print(f"rep {rep} path {path} early={os.environ.get('UMV_ATTN_EARLY','1')}: start->wait {(t[i,1]-t[i,0])/1e3:.1f}; after wait:", ", ".join(f"dbg{k}={(t[i,4+k]-t[i,1])/1e3:.1f}" for k in range(8)), f"first end {(t[i,2]-t[i,1])/1e3:.1f} last end {(t[i,3]-t[i,1])/1e3:.1f}")
