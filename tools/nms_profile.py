import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from abr_iod_b200.layers import nms_batched
from inputs import make_boxes
rng = np.random.default_rng(3)
for n,batch in ((6000,4),(12000,4)):
    data=[make_boxes(rng,n,1216,800) for _ in range(batch)]
    data=[(np.ascontiguousarray(b[np.argsort(-s, kind='stable')]), np.ascontiguousarray(np.sort(s)[::-1])) for b,s in data]  # RPN hands NMS a sorted top-k
    boxes=[torch.from_numpy(b).cuda() for b,_ in data]; scores=[torch.from_numpy(s).cuda() for _,s in data]
    for _ in range(2): k,c = nms_batched(boxes,scores,0.7,2000)
    torch.cuda.synchronize()
    print(n,batch,c.tolist())
