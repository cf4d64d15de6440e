R=128; RS=132
def A(r): return (r+3)&~3
def rot(r,XPL): return 4*(((r>>2) - XPL*r) & 7)
def S(r,XPL): return A(r)+rot(r,XPL)
def consumer_ok(XPL):
    for tau in range(128):
        for j in range(XPL):
            banks=set()
            for r in range(32):
                x=r*XPL+j
                col=(tau - r + S(r,XPL)) & 127
                b=(RS*x+col)%32
                if b in banks: return False
                banks.add(b)
    return True
print([consumer_ok(X) for X in range(1,9)])
# epilogue STS.128: thread = row x (consecutive), chunk c of tile t: col = (32t + S(r) + 4c)&127 ; quarter-warp = 8 consecutive x
def epi(XPL, tx):
    tot=0;n=0;worst=0
    for t in range(4):
      for c in range(8):
        for x0 in range(0, tx, 8):
            groups={}
            for x in range(x0,min(x0+8,tx)):
                r=x//XPL
                col=(32*t + S(r,XPL) + 4*c)&127
                g=((RS*x+col)//4)%8
                groups[g]=groups.get(g,0)+1
            m=max(groups.values()); tot+=m; n+=1; worst=max(worst,m)
    return round(tot/n,2), worst
print({X:epi(X,32*X) for X in range(1,9)})
