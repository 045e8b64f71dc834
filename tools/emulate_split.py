"""CPU emulation of the operand-split schemes considered for the tensor-core convolution (fp32 accumulate):
   TF32 hi/lo split vs fp16 hi/lo split with power-of-two weight scaling, both with 3 products, against the fp32 oracle.
   Result (6 seeded frames, 94 patches): max |dloc| 7.3e-4 vs 8.5e-4, max |dheat| 5.7e-7 vs 5.3e-7, 0 arg-max flips either way.
   Build-container tool (slow: float64 convolutions); not used by tests or the product."""
import sys, numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, __import__('os').path.join(__import__('os').path.dirname(__import__('os').path.abspath(__file__)), '..'))
import oracle
from deepcharuco_b200 import synth, weights_io as W
torch.set_num_threads(8)
sd, sr = W.load_state(W.DEFAULT_DEEPC), W.load_state(W.DEFAULT_REFINENET)

def split_tf32(x):
    b = x.contiguous().view(torch.int32)
    hi = ((b + 0x1000) & ~0x1fff).view(torch.float32)
    lo = x - hi
    bl = lo.contiguous().view(torch.int32)
    lo = ((bl + 0x1000) & ~0x1fff).view(torch.float32)
    return hi, lo
def split_f16(x, scale=1.0):
    xs = x*scale
    hi = xs.half().float()
    lo = (xs - hi).half().float()
    return hi, lo

def conv_split(x, w, b, pad, mode, wscale):
    if mode=='fp32':
        return F.conv2d(x, w, b, padding=pad)
    if mode=='tf32':
        xh,xl = split_tf32(x); wh,wl = split_tf32(w); s=1.0
    else:
        xh,xl = split_f16(x); wh,wl = split_f16(w, wscale); s=wscale
    main = F.conv2d(xh.double(), wh.double(), None, padding=pad)
    small = F.conv2d(xh.double(), wl.double(), None, padding=pad) + F.conv2d(xl.double(), wh.double(), None, padding=pad)
    y = (main.float() + small.float()) / s
    return y + b.view(1,-1,1,1)

def cbr(x, st, name, pad, mode):
    w = torch.from_numpy(st[name+'.weight']); b = torch.from_numpy(st[name+'.bias'])
    if w.shape[1] == 1:
        y = F.conv2d(x, w, b, padding=pad)
    else:
        mx = float(w.abs().max()); s = 2.0**np.floor(np.log2(32768.0/mx))
        y = conv_split(x, w, b, pad, mode, s)
    bn='bn'+name[4:]
    y = F.batch_norm(y, torch.from_numpy(st[bn+'.running_mean']), torch.from_numpy(st[bn+'.running_var']), torch.from_numpy(st[bn+'.weight']), torch.from_numpy(st[bn+'.bias']), False, 0.1, 1e-5)
    return F.relu(y)

def det(x, mode):
    st=sd
    x=cbr(x,st,'conv1a',1,mode); x=cbr(x,st,'conv1b',1,mode); x=F.max_pool2d(x,2,2)
    x=cbr(x,st,'conv2a',1,mode); x=cbr(x,st,'conv2b',1,mode); x=F.max_pool2d(x,2,2)
    x=cbr(x,st,'conv3a',1,mode); x=cbr(x,st,'conv3b',1,mode); x=F.max_pool2d(x,2,2)
    x=cbr(x,st,'conv4a',1,mode); x=cbr(x,st,'conv4b',1,mode)
    pa=cbr(x,st,'convPa',1,mode); da=cbr(x,st,'convDa',1,mode)
    loc=F.conv2d(pa, torch.from_numpy(st['convPb.weight']), torch.from_numpy(st['convPb.bias']))
    ids=F.conv2d(da, torch.from_numpy(st['convDb.weight']), torch.from_numpy(st['convDb.bias']))
    return loc, ids
def ref(x, mode):
    st=sr
    x=cbr(x,st,'conv1a',0,mode); x=cbr(x,st,'conv1b',0,mode); x=cbr(x,st,'conv2a',0,mode); x=cbr(x,st,'conv2b',0,mode)
    x=F.max_pool2d(x,2,2); x=cbr(x,st,'conv3a',1,mode); x=cbr(x,st,'conv3b',1,mode); x=F.interpolate(x,scale_factor=2,mode='nearest')
    x=cbr(x,st,'conv4a',1,mode); x=cbr(x,st,'conv4b',1,mode); x=F.interpolate(x,scale_factor=2,mode='nearest')
    x=cbr(x,st,'conv5a',1,mode); x=cbr(x,st,'conv5b',1,mode); x=F.interpolate(x,scale_factor=2,mode='nearest')
    x=cbr(x,st,'convPa',1,mode)
    return F.conv2d(x, torch.from_numpy(st['convPb.weight']), torch.from_numpy(st['convPb.bias']))

frames = synth.make_frames(6, seed=4)
with torch.no_grad():
    x = torch.from_numpy(np.stack([oracle.pre_bgr_image(f) for f in frames]))
    loc0, ids0 = oracle.detector_forward(sd, x)
    for mode in ('fp32','tf32','f16'):
        loc, ids = det(x, mode)
        print(mode, 'dloc', float((loc-loc0).abs().max()), 'dids', float((ids-ids0).abs().max()), 'argmax flips loc', int((loc.argmax(1)!=loc0.argmax(1)).sum()), 'ids', int((ids.argmax(1)!=ids0.argmax(1)).sum()))
    # refinenet on oracle patches
    P=[]; 
    for f in frames:
        r, st = oracle.pipeline.infer_gray(sd, sr, f, return_stages=True)
        P.append(st['patches'])
    P = torch.from_numpy(np.concatenate(P))[:,None]
    h0 = oracle.refinenet_forward(sr, P)
    for mode in ('fp32','tf32','f16'):
        h = ref(P, mode)
        fl = int((h.flatten(1).argmax(1)!=h0.flatten(1).argmax(1)).sum())
        print(mode, 'dheat', float((h-h0).abs().max()), 'flips', fl, 'of', h.shape[0])
    # activation ranges per layer (for fp16 range sanity)
    _,_,fd = oracle.detector_forward(sd, x, return_features=True)
    print('det act max', {k: round(float(v.max()),2) for k,v in fd.items()})
    _, fr = oracle.refinenet_forward(sr, P, return_features=True)
    print('ref act max', {k: round(float(v.max()),2) for k,v in fr.items()})
    print('w max det', {k[:-7]: round(float(np.abs(v).max()),3) for k,v in sd.items() if k.endswith('.weight') and k.startswith('conv')})
