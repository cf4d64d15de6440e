"""Build a variant of the library next to the product for same-box A/B runs (profiles/ab_variants.sh):
   python profiles/build_variant.py <dir with the .cu sources> <tag> [extra nvcc flags]  ->  art_tts_b200/lib/libmas_ab_<tag>.so
   e.g. the committed kernels:  git archive HEAD art_tts_b200/csrc include | tar -x -C /tmp/head && python profiles/build_variant.py /tmp/head/art_tts_b200/csrc B"""
import subprocess, glob, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from art_tts_b200 import build as b
def build_from(srcdir, out, extra=()):
    objs=[]; procs=[]
    os.makedirs('/tmp/abobj', exist_ok=True)
    for s in sorted(glob.glob(os.path.join(srcdir,'*.cu'))):
        o=os.path.join('/tmp/abobj', os.path.basename(out)+'_'+os.path.basename(s)+'.o')
        objs.append(o)
        procs.append(subprocess.Popen(['nvcc',*b.NVCC_FLAGS,*extra,'-c','-o',o,s],stdout=subprocess.DEVNULL,stderr=subprocess.DEVNULL))
    for p in procs: assert p.wait()==0
    subprocess.check_call(['nvcc','-shared','-gencode','arch=compute_100a,code=sm_100a','-o',out,*objs])
if __name__ == '__main__':
    build_from(sys.argv[1], os.path.join(b.LIBDIR, 'libmas_ab_%s.so' % sys.argv[2]), sys.argv[3:])
