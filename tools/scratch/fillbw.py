import torch, time
x = torch.empty(512**3, dtype=torch.float32, device='cuda')
y = torch.empty(1024**3, dtype=torch.float32, device='cuda')
for t, name in ((x, '512^3'), (y, '1024^3')):
    for _ in range(3): t.fill_(1.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(10):
        e0.record(); t.fill_(1.0); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts)//2]
    print(name, 'fill_', ms, 'ms', t.numel()*4/ms/1e6, 'GB/s')
    ts = []
    for _ in range(10):
        e0.record(); torch.cuda.current_stream(); 
        import ctypes
        t.zero_(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts)//2]
    print(name, 'zero_ (memset)', ms, 'ms', t.numel()*4/ms/1e6, 'GB/s')
