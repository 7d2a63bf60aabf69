"""fm_device.cuh: extend_one_text_packed (SwDriver::extend of a unique seed hit, 32 characters per step over the 2-bit packed read and
the 2-bit packed joined text) compiled FOR THE HOST and fuzzed against the per-character loop it replaces (extend_one_text's body):
both strands, both directions, Ns in the read, the two ends of the text (the "$" row), garbage beyond the read's and the text's last
word.  The device build of the same source is covered by tests/test_extend.py and tests/test_xengine_gpu.py on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS_HEAD = r'''
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
static inline uint64_t __brevll(uint64_t x){uint64_t r=0;for(int i=0;i<64;i++){r=(r<<1)|(x&1);x>>=1;}return r;}
static inline uint32_t __brev(uint32_t x){uint32_t r=0;for(int i=0;i<32;i++){r=(r<<1)|(x&1);x>>=1;}return r;}
static inline int __ffsll(long long x){return x?__builtin_ctzll((unsigned long long)x)+1:0;}
struct Fw{uint64_t len;}; template<typename OFF> struct DevIndex{Fw fw; const uint8_t*refBuf;};
'''

HARNESS_TAIL = r'''
static int read_char(const uint8_t*seq,int len,int strand,int pos){ if(strand==0) return seq[pos]; int c=seq[len-1-pos]; return c>3?4:3-c; }
static uint32_t ref_ext(const DevIndex<uint32_t>&ix,int64_t b,int tstep,const uint8_t*s,int len,int strand,int i0,int step,int lim){
  uint32_t n=0; int64_t tlen=ix.fw.len;
  for(int ii=0;ii<lim;ii++){ int rdc=read_char(s,len,strand,i0+ii*step); int c=-1; if(b>=0&&b<tlen){c=(ix.refBuf[b>>2]>>((b&3)<<1))&3; b+=tstep;} if(c!=rdc&&rdc<=3)break; if(++n==255)break;} return n; }
int main(){ srand(5); long bad=0,tot=0;
 for(int it=0;it<400000;it++){
   int tlen=1+rand()%600; std::vector<uint8_t> T(tlen); for(auto&c:T)c=rand()%4;
   std::vector<uint8_t> buf((tlen+3)/4+16,0); for(int i=0;i<tlen;i++) buf[i>>2]|=T[i]<<((i&3)*2);
   for(int i=tlen;i<((tlen+3)/4+8)*4;i++) buf[i>>2]|=(rand()%4)<<((i&3)*2);
   std::vector<uint64_t> bufw((buf.size()+7)/8+1); memcpy(bufw.data(),buf.data(),buf.size());
   DevIndex<uint32_t> ix; ix.fw.len=tlen; ix.refBuf=(const uint8_t*)bufw.data();
   int len=1+rand()%300; std::vector<uint8_t> rd(len);
   int strand=rand()%2;
   int anchor=rand()%tlen - 50;
   for(int i=0;i<len;i++){ int64_t tp=anchor+i; int c=(tp>=0&&tp<tlen)?T[tp]:rand()%4; if(rand()%40==0)c=rand()%4; if(rand()%50==0)c=4; rd[i]=c; }
   std::vector<uint8_t> raw(len); if(strand==0) raw=rd; else for(int i=0;i<len;i++){int c=rd[len-1-i]; raw[i]=c>3?4:3-c;}
   std::vector<uint64_t> pk((len+31)/32+2,0); std::vector<uint32_t> nm((len+31)/32+2,0);
   for(int i=0;i<len;i++){ if(raw[i]>3) nm[i>>5]|=1u<<(i&31); else pk[i>>5]|=(uint64_t)raw[i]<<(2*(i&31)); }
   for(int i=len;i<(int)pk.size()*32;i++){ pk[i>>5]|=(uint64_t)(rand()%4)<<(2*(i&31)); if(rand()%3==0) nm[i>>5]|=1u<<(i&31); }
   int dir=rand()%2; int step=dir?1:-1, tstep=step;
   int i0=rand()%len; int lim= step>0 ? len-i0 : i0+1; if(lim>1&&rand()%3==0) lim=1+rand()%lim;
   int64_t b = anchor + i0 + (rand()%5==0 ? (rand()%7-3):0);
   if(rand()%10==0) b = step>0 ? tlen-1-rand()%40 : rand()%40;
   if(step>0){ if(b<0)b=0; if(b>tlen)b=tlen; } else { if(b<-1)b=-1; if(b>tlen-1)b=tlen-1; }     // the starts SwDriver::extend can produce
   uint32_t want=ref_ext(ix,b,tstep,raw.data(),len,strand,i0,step,lim);
   uint32_t got=extend_one_text_packed<uint32_t>(ix,b,tstep,pk.data(),nm.data(),len,strand,i0,step,lim);
   tot++; if(want!=got){ if(bad<5) printf("MISMATCH tlen=%d len=%d strand=%d step=%d i0=%d lim=%d b=%ld want=%u got=%u\n",tlen,len,strand,step,i0,lim,(long)b,want,got); bad++; }
 }
 printf("%ld cases, %ld bad\n",tot,bad); return bad!=0; }
'''


@pytest.mark.timeout(300)
def test_word_parallel_text_extension_equals_the_per_character_walk(tmp_path):
    src = open(os.path.join(ROOT, "bowtie2_b200", "csrc", "fm_device.cuh")).read()
    a = src.index("__device__ __forceinline__ uint64_t swap_rev_pairs")
    b = src.index("// SwDriver::extend, both directions of one seed hit")
    code = src[a:b].replace("__device__ __forceinline__", "static inline").replace("__ldg(tw + k)", "tw[k]")
    cpp = tmp_path / "t.cpp"
    cpp.write_text(HARNESS_HEAD + code + HARNESS_TAIL)
    exe = str(tmp_path / "t")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, str(cpp)])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "400000 cases, 0 bad" in out.stdout, out.stdout[-600:]
