import numpy as np
from scipy.special import erfc, log_ndtr
from scipy.optimize import linprog
def gelu(x): return x*0.5*erfc(-x/np.sqrt(2))
def approx(c, x, A):
    a = np.minimum(np.abs(x), A)
    q = np.zeros_like(x)
    for k in c[::-1]: q = q*a + k
    return np.maximum(x,0) - np.abs(x)*np.exp2(-q)
for A in (6.0, 8.0):
  a = np.linspace(0, A, 6001)
  Q = -(log_ndtr(-a))/np.log(2)      # = -log2(Phi(-a)); includes the 0.5 -> Q(0)=1
  w = a*np.exp2(-Q)*np.log(2)
  for deg in (4,5,6,7):
    V = np.vander(a, deg+1, increasing=True)
    n = deg+1
    Am = np.vstack([np.hstack([V*w[:,None], -np.ones((len(a),1))]), np.hstack([-V*w[:,None], -np.ones((len(a),1))])])
    b = np.concatenate([w*Q, -w*Q])
    cost = np.zeros(n+1); cost[-1]=1
    r = linprog(cost, A_ub=Am, b_ub=b, bounds=[(None,None)]*n+[(0,None)], method='highs')
    c = r.x[:n]
    xx = np.linspace(-12, 12, 200001)
    e = np.abs(approx(c, xx, A)-gelu(xx))
    # fp32 evaluation
    x32 = xx.astype(np.float32); a32 = np.minimum(np.abs(x32), np.float32(A)); q = np.zeros_like(x32)
    for k in c[::-1]: q = q*a32 + np.float32(k)
    y32 = np.maximum(x32,0) - np.abs(x32)*np.exp2(-q)
    e32 = np.abs(y32.astype(np.float64)-gelu(xx))
    print(A, deg, ["%.9e"%v for v in c], "lin t %.3e"%r.x[-1], "maxabs %.3e at %.3f"%(e.max(), xx[e.argmax()]), "fp32 max %.3e rms %.3e"%(e32.max(), np.sqrt((e32[np.abs(xx)<4]**2).mean())))
