#!/usr/bin/env python
"""Recovers cv::ORB's rBRIEF test-location table (bit_pattern_31_, 256 tests x 2 points) from the installed cv2 as a black
box (OpenCV's sources are not in this image).  One key point at angle 0 in the centre of a synthetic image; the image is a
step edge at position c along x, y, x+y or x-y, in both polarities.  ORB blurs the image (7x7, sigma 2) and sets bit i when
blurred(p0) < blurred(p1); sweeping c and comparing with the same blur applied here determines (x0, x1), (y0, y1) and, for
the tests with x0 == x1 or y0 == y1, the diagonals disambiguate.  Writes /tmp/exp/orb_pattern.npy; the .inc files under
oracle/ and vi-slam_b200/csrc/ were generated from it.  Ends with a validation against cv2 on random images and angles."""
import os
os.makedirs('/tmp/exp', exist_ok=True)
import cv2, numpy as np, itertools
S=160; C=80
orb=cv2.ORB_create(nfeatures=10,nlevels=1,edgeThreshold=31,patchSize=31,fastThreshold=20)
kp=[cv2.KeyPoint(float(C),float(C),31,0.0,0,0,-1)]
yy,xx=np.mgrid[0:S,0:S]; X=xx-C; Y=yy-C
def bits(img):
    k,d=orb.compute(img,kp)
    assert len(k)==1
    return np.unpackbits(d[0],bitorder='little')   # bit i of byte j = test 8*j+i
R=range(-26,27)
def probe(F):
    obs=[];Bs=[]
    for c in R:
        for pol in (0,1):
            I=np.where(F>=c,255,0).astype(np.uint8)
            if pol: I=255-I
            obs.append(bits(I)); Bs.append(cv2.GaussianBlur(I,(7,7),2,2,borderType=cv2.BORDER_REFLECT_101).astype(np.int32))
    return np.array(obs),Bs
# x and y probes
ox,Bx=probe(X); oy,By=probe(Y)
rng=range(-15,16)
def solve1d(obs,Bs,axis):
    # value along the axis through the centre
    prof=np.array([ (B[C,C-15:C+16] if axis==0 else B[C-15:C+16,C]) for B in Bs])   # [nprobe,31]
    pred=prof[:,:,None]<prof[:,None,:]      # [nprobe,31(p0),31(p1)]
    sols=[]
    for b in range(256):
        m=(pred==obs[:,b][:,None,None].astype(bool)).all(axis=0)
        sols.append(np.argwhere(m)-15)
    return sols
sx=solve1d(ox,Bx,0); sy=solve1d(oy,By,1)
print("x unique",sum(len(s)==1 for s in sx),"ambig",sum(len(s)>1 for s in sx),"none",sum(len(s)==0 for s in sx))
print("y unique",sum(len(s)==1 for s in sy),"ambig",sum(len(s)>1 for s in sy),"none",sum(len(s)==0 for s in sy))
np.save('/tmp/exp/sx.npy',np.array([s[0] if len(s)==1 else [99,99] for s in sx])); np.save('/tmp/exp/sy.npy',np.array([s[0] if len(s)==1 else [99,99] for s in sy]))
amb=[b for b in range(256) if len(sx[b])!=1 or len(sy[b])!=1]
print("ambiguous bits",amb[:40], [ (len(sx[b]),len(sy[b])) for b in amb[:10]])
od,Bd=probe(X+Y); oa,Ba=probe(X-Y)
pat=np.zeros((256,2,2),np.int32)   # [bit][point][x,y]
bad=0
for b in range(256):
    cx=[tuple(s) for s in sx[b]]; cy=[tuple(s) for s in sy[b]]
    good=[]
    for (x0,x1) in cx:
        for (y0,y1) in cy:
            ok=True
            for obs,Bs in ((od,Bd),(oa,Ba)):
                for i,B in enumerate(Bs):
                    if (B[C+y0,C+x0]<B[C+y1,C+x1])!=bool(obs[i][b]): ok=False;break
                if not ok: break
            if ok: good.append((x0,y0,x1,y1))
    if len(good)!=1: bad+=1; print("bit",b,"solutions",len(good)); continue
    x0,y0,x1,y1=good[0]; pat[b]=[[x0,y0],[x1,y1]]
print("unresolved",bad); print(pat[:4].reshape(4,4), pat.min(), pat.max())
np.save('/tmp/exp/orb_pattern.npy',pat)
# validate on random images, angle 0 and non-zero angles
rngm=np.random.default_rng(1)
def desc_sim(img,x,y,ang):
    B=cv2.GaussianBlur(img,(7,7),2,2,borderType=cv2.BORDER_REFLECT_101).astype(np.int32)
    a=np.float32(np.cos(np.float32(ang)*np.float32(np.pi/180))); bb=np.float32(np.sin(np.float32(ang)*np.float32(np.pi/180)))
    out=np.zeros(256,np.uint8)
    for i in range(256):
        v=[]
        for p in range(2):
            px,py=np.float32(pat[i,p,0]),np.float32(pat[i,p,1])
            fx=np.float32(np.float32(px*a)-np.float32(py*bb)); fy=np.float32(np.float32(px*bb)+np.float32(py*a))
            ix=int(np.rint(fx)); iy=int(np.rint(fy))
            v.append(B[y+iy,x+ix])
        out[i]=v[0]<v[1]
    return out
tot=0;okc=0
for t in range(20):
    img=cv2.GaussianBlur((rngm.random((S,S))*255).astype(np.uint8),(5,5),1.5)
    ang=float(rngm.uniform(0,360)) if t>4 else 0.0
    k=[cv2.KeyPoint(float(C),float(C),31,ang,0,0,-1)]
    _,d=orb.compute(img,k)
    got=np.unpackbits(d[0],bitorder='little'); sim=desc_sim(img,C,C,ang)
    tot+=256; okc+=(got==sim).sum()
    if (got!=sim).any(): print("t",t,"ang",ang,"mismatch bits",np.nonzero(got!=sim)[0][:10])
print("validation",okc,"/",tot)
