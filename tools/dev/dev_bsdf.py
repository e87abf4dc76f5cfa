import sys, numpy as np
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import orc, cudapath
ctx=cudapath.Context(0); osc=orc.Scene()
HAIR=(0.143016, 0.0156076, 1.80928e-005)
mats=[('kajiyakay', dict(diffuseReflectance=HAIR, exponent=10.0)),('marschner', dict(intIOR=1.55, extIOR=1.0, specularReflectance=(0.592384, 0.32628, 0.0528657))),('marschner', dict(intIOR=1.55, extIOR=1.0, alpha=0.2, distribution='ggx', diffuseReflectance=HAIR))]
for t,p in mats: ctx.add_bsdf(t,**p); osc.add_bsdf(t,**p)
ctx.add_hair(np.array([[0,0,0],[0,1,0],[0.1,2,0]],np.float32), np.array([1,0,0],np.uint8), 0.05, 0)
ctx.set_camera(np.eye(4,dtype=np.float32),35.0,width=16,height=16); ctx.build()
rng=np.random.default_rng(1)
def sph(n):
    v=rng.normal(size=(n,3)); v/=np.linalg.norm(v,axis=1,keepdims=True); return v.astype(np.float32)
n=1<<20
wi,wo=sph(n),sph(n)
for b in range(3):
    ge,gp=ctx.bsdf_eval(b,wi,wo); oe,op=osc.bsdf_eval(b,wi,wo)
    sc=np.abs(oe).max(); err=np.abs(ge-oe)/np.maximum(np.abs(oe),1e-6*sc)
    print('eval bsdf',b,'max',err.max(),'q99.99',np.quantile(err,0.9999),'n>1e-5',(err>1e-5).sum(),'exact frac',(ge==oe).mean())
smp=rng.random((n,2)).astype(np.float32)
for b in range(3):
    g=ctx.bsdf_sample(b,wi,smp); o=osc.bsdf_sample(b,wi,smp)
    same=g[3]==o[3]; v=same&(np.abs(o[1]).sum(1)>0)
    sc=np.abs(o[1][v]).max(); err=np.abs(g[1][v]-o[1][v])/np.maximum(np.abs(o[1][v]),1e-6*sc)
    print('sample bsdf',b,'mismatch',(~same).sum(),'wo maxdiff',np.abs(g[0][v]-o[0][v]).max(),'w err max',err.max(),'q99.9',np.quantile(err,0.999),'exact',(g[1][v]==o[1][v]).mean())
