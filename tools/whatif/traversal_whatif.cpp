// What-if study on the CPU (no GPU, not part of the product or of the oracle): how many loop iterations (V) and leaf
// visits (L) of intersectScene (tracer.fs:366-404) would (1) pop-time pruning -- deferred nodes re-tested against the
// current hit distance when popped, which the reference does not do (tracer.fs:401) -- and (2) a 4-wide collapse of the
// same binary tree (grandchildren, sorted near-first, pop-time pruning) save?  Run by tools/whatif/run.py on the bench
// scenes; primary rays, cosine-distributed bounce rays from the hit points and hit-or-miss rays towards a jittered sun.
// Results are quoted in DESIGN.md section 9.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <random>
#include <algorithm>
using namespace std;
static vector<float> bvh, tris; int N, T;
const float MAXT=100000.0f, EPS=1e-6f;
static inline int ib(int node,int k){int v; memcpy(&v,&bvh[(size_t)node*9+k],4); return v;}
static float slab(const float* b, const float* o, const float* inv){
  float t1x=(b[0]-o[0])*inv[0], t2x=(b[3]-o[0])*inv[0];
  float t1y=(b[1]-o[1])*inv[1], t2y=(b[4]-o[1])*inv[1];
  float t1z=(b[2]-o[2])*inv[2], t2z=(b[5]-o[2])*inv[2];
  float tMax=fminf(fminf(fmaxf(t1x,t2x),fmaxf(t1y,t2y)),fmaxf(t1z,t2z));
  float tMin=fmaxf(fmaxf(fminf(t1x,t2x),fminf(t1y,t2y)),fminf(t1z,t2z));
  return (tMax>=tMin && tMax>0)? tMin: MAXT;
}
static float tri(int t,const float*o,const float*d){
  if(t>=T) return MAXT;
  const float* v=&tris[(size_t)t*9];
  float e1[3]={v[3]-v[0],v[4]-v[1],v[5]-v[2]}, e2[3]={v[6]-v[0],v[7]-v[1],v[8]-v[2]};
  float p[3]={d[1]*e2[2]-e2[1]*d[2], d[2]*e2[0]-e2[2]*d[0], d[0]*e2[1]-e2[0]*d[1]};
  float det=e1[0]*p[0]+e1[1]*p[1]+e1[2]*p[2];
  if(fabsf(det)<EPS) return MAXT;
  float inv=1.0f/det; float tv[3]={o[0]-v[0],o[1]-v[1],o[2]-v[2]};
  float u=(tv[0]*p[0]+tv[1]*p[1]+tv[2]*p[2])*inv; if(u<0||u>1) return MAXT;
  float q[3]={tv[1]*e1[2]-e1[1]*tv[2], tv[2]*e1[0]-e1[2]*tv[0], tv[0]*e1[1]-e1[0]*tv[1]};
  float w=(d[0]*q[0]+d[1]*q[1]+d[2]*q[2])*inv; if(w<0||u+w>1) return MAXT;
  float dist=(e2[0]*q[0]+e2[1]*q[1]+e2[2]*q[2])*inv; return dist>EPS?dist:MAXT;
}
struct Res{int idx; float t; long V,L;};
// mode 0 reference, 1 pop-prune
static Res trace(const float*o,const float*d,int mode,bool anyhit){
  float inv[3]={1.0f/d[0],1.0f/d[1],1.0f/d[2]};
  Res r{-1,MAXT,0,0};
  int stack[64]; float st[64]; int sp=0; stack[sp]=-1; st[sp]=0; sp++;
  int cur=0;
  while(cur!=-1){
    r.V++;
    int triIdx=ib(cur,2);
    if(triIdx>-1){
      r.L++;
      for(int k=0;k<4;k++){float x=tri(triIdx+k,o,d); if(x<r.t){r.t=x;r.idx=triIdx+k;}}
      if(anyhit && r.idx!=-1) return r;
      // pop
      for(;;){ sp--; cur=stack[sp]; if(cur==-1||mode==0||st[sp]<r.t) break; }
      continue;
    }
    int l=ib(cur,0), rr=ib(cur,1);
    float lh=slab(&bvh[(size_t)l*9+3],o,inv), rh=slab(&bvh[(size_t)rr*9+3],o,inv);
    bool tl=lh<r.t, tr=rh<r.t;
    if(tl&&tr){ bool rf=lh>rh; stack[sp]=rf?l:rr; st[sp]=rf?lh:rh; sp++; cur=rf?rr:l; }
    else if(tl||tr) cur=tl?l:rr;
    else { for(;;){ sp--; cur=stack[sp]; if(cur==-1||mode==0||st[sp]<r.t) break; } }
  }
  return r;
}
// mode 2: 4-wide collapse (grandchildren), sorted near-first, pop-time pruning
static Res trace4(const float*o,const float*d,bool anyhit){
  float inv[3]={1.0f/d[0],1.0f/d[1],1.0f/d[2]};
  Res r{-1,MAXT,0,0};
  int stack[128]; float st[128]; int sp=0; stack[sp]=-1; st[sp]=0; sp++;
  int cur=0;
  while(cur!=-1){
    int triIdx=ib(cur,2);
    if(triIdx>-1){
      r.L++;
      for(int k=0;k<4;k++){float x=tri(triIdx+k,o,d); if(x<r.t){r.t=x;r.idx=triIdx+k;}}
      if(anyhit && r.idx!=-1) return r;
    } else {
      r.V++;
      int ch[4]; int nc=0; int l=ib(cur,0), rr=ib(cur,1);
      int two[2]={l,rr};
      for(int k=0;k<2;k++){ int c=two[k]; if(ib(c,2)>-1) ch[nc++]=c; else { ch[nc++]=ib(c,0); ch[nc++]=ib(c,1);} }
      float h[4]; int id[4]; int nh=0;
      for(int k=0;k<nc;k++){ float t=slab(&bvh[(size_t)ch[k]*9+3],o,inv); if(t<r.t){ h[nh]=t; id[nh]=ch[k]; nh++; } }
      // sort descending by t (farthest first pushed)
      for(int a=0;a<nh;a++)for(int b=a+1;b<nh;b++) if(h[b]>h[a]){swap(h[a],h[b]);swap(id[a],id[b]);}
      for(int a=0;a<nh;a++){ stack[sp]=id[a]; st[sp]=h[a]; sp++; }
    }
    for(;;){ sp--; cur=stack[sp]; if(cur==-1||st[sp]<r.t) break; }
  }
  return r;
}
int main(int argc,char**argv){
  const char* name=argv[1];
  char fn[256]; snprintf(fn,256,"%s_bvh.bin",name); FILE*f=fopen(fn,"rb"); fseek(f,0,SEEK_END); long sz=ftell(f); fseek(f,0,SEEK_SET); bvh.resize(sz/4); fread(bvh.data(),1,sz,f); fclose(f); N=sz/36;
  snprintf(fn,256,"%s_tris.bin",name); f=fopen(fn,"rb"); fseek(f,0,SEEK_END); sz=ftell(f); fseek(f,0,SEEK_SET); tris.resize(sz/4); fread(tris.data(),1,sz,f); fclose(f); T=sz/36;
  float eye[3]={(float)atof(argv[2]),(float)atof(argv[3]),(float)atof(argv[4])}, dir[3]={(float)atof(argv[5]),(float)atof(argv[6]),(float)atof(argv[7])};
  // camera basis
  float bx[3]={dir[1]*0-1*dir[2]*0+(-dir[2]), 0, dir[0]}; // cross(I,(0,1,0)) = (-Iz,0,Ix)
  bx[0]=-dir[2]; bx[1]=0; bx[2]=dir[0]; float n=sqrtf(bx[0]*bx[0]+bx[2]*bx[2]); bx[0]/=n; bx[2]/=n;
  float by[3]={bx[1]*dir[2]-dir[1]*bx[2], bx[2]*dir[0]-dir[2]*bx[0], bx[0]*dir[1]-dir[0]*bx[1]}; n=sqrtf(by[0]*by[0]+by[1]*by[1]+by[2]*by[2]); for(auto&x:by)x/=n;
  mt19937 rng(1); uniform_real_distribution<float> U(0,1);
  const int W=320,H=180;
  long V[3][3]={{0}},L[3][3]={{0}},cnt[3]={0}; long mism=0, mism4=0;
  for(int y=0;y<H;y++)for(int x=0;x<W;x++){
    float u=((x+0.5f)/W*2-1)*0.5f*(float)W/H, v=((y+0.5f)/H*2-1)*0.5f;
    float d[3]; for(int k=0;k<3;k++) d[k]=dir[k]+bx[k]*u+by[k]*v; n=sqrtf(d[0]*d[0]+d[1]*d[1]+d[2]*d[2]); for(auto&q:d)q/=n;
    float o[3]={eye[0],eye[1],eye[2]};
    for(int b=0;b<4;b++){
      Res r0=trace(o,d,0,false), r1=trace(o,d,1,false);
      if(r0.idx!=r1.idx||r0.t!=r1.t) mism++; Res r2=trace4(o,d,false); if(r0.idx!=r2.idx||r0.t!=r2.t) mism4++; V[2][b==0?0:1]+=r2.V; L[2][b==0?0:1]+=r2.L;
      int c=b==0?0:1; V[0][c]+=r0.V;L[0][c]+=r0.L;V[1][c]+=r1.V;L[1][c]+=r1.L;cnt[c]++;
      if(r0.idx<0) break;
      // hit point, normal
      const float* tv=&tris[(size_t)r0.idx*9]; float e1[3]={tv[3]-tv[0],tv[4]-tv[1],tv[5]-tv[2]},e2[3]={tv[6]-tv[0],tv[7]-tv[1],tv[8]-tv[2]};
      float nn[3]={e1[1]*e2[2]-e1[2]*e2[1],e1[2]*e2[0]-e1[0]*e2[2],e1[0]*e2[1]-e1[1]*e2[0]}; n=sqrtf(nn[0]*nn[0]+nn[1]*nn[1]+nn[2]*nn[2]); if(n==0)break; for(auto&q:nn)q/=n;
      if(nn[0]*d[0]+nn[1]*d[1]+nn[2]*d[2]>0) for(auto&q:nn)q=-q;
      for(int k=0;k<3;k++) o[k]=o[k]+d[k]*r0.t+nn[k]*2e-6f;
      // shadow-like any-hit ray toward random upper-hemisphere dir ("sun" dir jittered)
      float sd[3]={0.3f+0.1f*U(rng),0.8f,0.4f+0.1f*U(rng)}; n=sqrtf(sd[0]*sd[0]+sd[1]*sd[1]+sd[2]*sd[2]); for(auto&q:sd)q/=n;
      Res s0=trace(o,sd,0,true), s1=trace(o,sd,1,true); V[0][2]+=s0.V;L[0][2]+=s0.L;V[1][2]+=s1.V;L[1][2]+=s1.L;cnt[2]++; Res s2=trace4(o,sd,true); V[2][2]+=s2.V; L[2][2]+=s2.L;
      // cosine-ish bounce
      float r1_=U(rng), r2_=U(rng); float ph=6.2831853f*r2_, rr=sqrtf(r1_); float lx=rr*cosf(ph), ly=rr*sinf(ph), lz=sqrtf(fmaxf(0.f,1-lx*lx-ly*ly));
      float up[3]={0,0,1}; if(fabsf(nn[2])>=0.999f){up[0]=1;up[2]=0;}
      float tg[3]={up[1]*nn[2]-up[2]*nn[1],up[2]*nn[0]-up[0]*nn[2],up[0]*nn[1]-up[1]*nn[0]}; n=sqrtf(tg[0]*tg[0]+tg[1]*tg[1]+tg[2]*tg[2]); for(auto&q:tg)q/=n;
      float bt[3]={nn[1]*tg[2]-nn[2]*tg[1],nn[2]*tg[0]-nn[0]*tg[2],nn[0]*tg[1]-nn[1]*tg[0]};
      for(int k=0;k<3;k++) d[k]=tg[k]*lx+bt[k]*ly+nn[k]*lz;
    }
  }
  const char* nm[3]={"primary","bounce(closest)","shadow(anyhit)"};
  for(int c=0;c<3;c++) printf("%-16s rays %8ld | ref V %.2f L %.2f | pop-prune V %.2f L %.2f | wide4 steps %.2f L %.2f\n",nm[c],cnt[c],(double)V[0][c]/cnt[c],(double)L[0][c]/cnt[c],(double)V[1][c]/cnt[c],(double)L[1][c]/cnt[c],(double)V[2][c]/cnt[c],(double)L[2][c]/cnt[c]);
  printf("closest-hit mismatches (index or t): pop-prune %ld, wide4 %ld\n",mism,mism4);
}
