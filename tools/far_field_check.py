"""Accuracy of the far-field treatment of the far wings (k_far_nodes, csrc/sr_voigt.cu): Chebyshev
interpolation of the region-1 rational over a 512-point tile from n nodes, for lines at least
dfac tile lengths (+ ry + 1 Doppler widths) away; worst relative error per line over ry, xs, wing."""
import numpy as np
from numpy.polynomial import chebyshev as C
rng=np.random.default_rng(1)
TP=512
def K1(x,ry):
    u=x*x+ry*ry-0.5
    return 0.5641896*ry*(u+1)/(u*u+2*ry*ry)
def test(n_nodes, dfac, ry, xs, side=1, ntest=2000):
    # tile points P=0..511 ; line at distance: nearest tile point at x = d (in Doppler units), far side d+L
    L=(TP-1)*xs
    worst=0
    for _ in range(ntest):
        d=(dfac*L+ry+1)*(1+rng.uniform(0,3)**3) if rng.uniform()<0.7 else dfac*L+ry+1
        P=np.arange(TP)
        x=(d+P*xs) if side>0 else -(d+(TP-1-P)*xs)
        f=K1(x,ry)
        # chebyshev nodes on [0, TP-1]
        k=np.arange(n_nodes)
        t=np.cos((2*k+1)*np.pi/(2*n_nodes))
        Pn=0.5*(TP-1)*(1+t)
        xn=(d+Pn*xs) if side>0 else -(d+(TP-1-Pn)*xs)
        fn=K1(xn,ry)
        coef=C.chebfit(t,fn,n_nodes-1)
        fi=C.chebval(2*P/(TP-1)-1,coef)
        worst=max(worst,np.max(np.abs(fi-f)/np.abs(f)))
    return worst
for n_nodes in (8,10,12,16):
    for dfac in (1,2,3,4):
        w=max(test(n_nodes,dfac,ry,xs,side) for ry in (1e-4,0.05,1.0,10.0,80.0) for xs in (0.02,0.12,0.5) for side in (1,-1))
        print("nodes",n_nodes,"d >=",dfac,"L : worst rel err per line %.2e"%w)

print("--- monomial (Horner) evaluation, 12 nodes, d >= 2L")
def test_mono(n_nodes, dfac, ry, xs, side=1, ntest=500):
    L=(TP-1)*xs; worst=0
    k=np.arange(n_nodes); t=np.cos((2*k+1)*np.pi/(2*n_nodes))
    # transform matrix nodes -> cheb coef -> monomial coef
    Tm=np.array([[ (1.0 if j==0 else 2.0)/n_nodes*np.cos(j*(2*n+1)*np.pi/(2*n_nodes)) for n in range(n_nodes)] for j in range(n_nodes)])
    C2P=np.zeros((n_nodes,n_nodes))
    for j in range(n_nodes):
        e=np.zeros(n_nodes); e[j]=1; p=C.cheb2poly(e); C2P[:len(p),j]=p
    M=C2P@Tm   # nodes -> monomial coefs
    for _ in range(ntest):
        d=(dfac*L+ry+1)*(1+rng.uniform(0,3)**3) if rng.uniform()<0.7 else dfac*L+ry+1
        P=np.arange(TP); x=(d+P*xs) if side>0 else -(d+(TP-1-P)*xs)
        f=K1(x,ry)
        Pn=0.5*(TP-1)*(1+t); xn=(d+Pn*xs) if side>0 else -(d+(TP-1-Pn)*xs)
        m=M@K1(xn,ry)
        tt=2*P/(TP-1)-1
        fi=np.zeros(TP)
        for c in m[::-1]: fi=fi*tt+c
        worst=max(worst,np.max(np.abs(fi-f)/np.abs(f)))
    return worst
w=max(test_mono(12,2,ry,xs,side) for ry in (1e-4,0.05,1.0,10.0,80.0) for xs in (0.02,0.12,0.5) for side in (1,-1))
print("worst", w)
