// Corrupt valid BGZF payloads and run the decoder under ASan/UBSan: it must decline or decode, never touch
// memory outside [in, in+in_len+8) and [out, out+out_len)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstring>
#include <cstdint>
#include <random>
#include "inflate_fast.hpp"
using namespace inqhost;
static uint16_t rd16(const uint8_t*p){return p[0]|(p[1]<<8);} static uint32_t rd32(const uint8_t*p){return p[0]|(p[1]<<8)|(p[2]<<16)|((uint32_t)p[3]<<24);}
int main(int argc,char**argv){
  FILE*fp=fopen(argv[1],"rb"); std::vector<uint8_t> file; uint8_t buf[1<<16]; size_t k; while((k=fread(buf,1,sizeof buf,fp))>0) file.insert(file.end(),buf,buf+k); fclose(fp); file.resize(file.size()+16);
  int iters=argc>2?atoi(argv[2]):1000; std::mt19937_64 rng(12345);
  struct Blk{size_t off,in_len;uint32_t isize;}; std::vector<Blk> blks; size_t p=0;
  while(p+28<=file.size()-16 && blks.size()<400){ uint16_t xlen=rd16(&file[p+10]); size_t bsize=(size_t)rd16(&file[p+16])+1; uint32_t isize=rd32(&file[p+bsize-4]); if(isize) blks.push_back({p+12+xlen,bsize-12-xlen-8,isize}); p+=bsize; }
  FastInflater fi; size_t ok=0, declined=0;
  for(int it=0;it<iters;++it){ const Blk&b=blks[rng()%blks.size()];
    // exact-size heap copies so that ASan sees any overrun: input gets its 8 readable trailer bytes
    std::vector<uint8_t> in(file.begin()+b.off, file.begin()+b.off+b.in_len+8);
    size_t out_len=b.isize; int mode=rng()%6;
    if(mode==0){ int n=1+rng()%4; for(int j=0;j<n;++j) in[rng()%b.in_len]^=(uint8_t)(1u<<(rng()%8)); }
    else if(mode==1){ size_t a=rng()%b.in_len, n=1+rng()%64; for(size_t j=a;j<std::min(b.in_len,a+n);++j) in[j]=(uint8_t)rng(); }
    else if(mode==2){ out_len = rng()%(b.isize+1); }                 // output smaller than the stream wants
    else if(mode==3){ size_t cut=rng()%b.in_len; in.resize(cut+8); }  // truncated input
    else if(mode==4){ for(int j=0;j<8;++j) in[rng()%std::min<size_t>(b.in_len,40)]=(uint8_t)rng(); }   // header damage
    else { out_len = b.isize + 1 + rng()%100; }
    size_t in_len = in.size()-8;
    std::vector<uint8_t> out(out_len ? out_len : 1);
    if(fi.inflate(in.data(), in_len, out.data(), out_len)) ++ok; else ++declined; }
  printf("fuzz: %d cases, %zu decoded, %zu declined\n", iters, ok, declined);
}
