// CPU study: node visits / triangle tests per ray of the shipped 4-wide traversal order (children sorted by
// entry distance, popped entries culled by the current best) for different BVH2 -> BVH4 collapse strategies.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>
#include <string>
#include "../include/odinrt_b200.h"
struct WNode { float lo[4][3], hi[4][3]; int32_t child[4]; };  // child>=0 inner, <0 leaf ~((first<<3)|cnt), EMPTY
static const int32_t EMPTY = (int32_t)0x80000000;
static std::vector<char> slurp(const char* p){FILE*f=fopen(p,"rb");fseek(f,0,SEEK_END);long n=ftell(f);fseek(f,0,SEEK_SET);std::vector<char> b(n);fread(b.data(),1,n,f);fclose(f);return b;}
static const ort_bvh_node* B; static int64_t NB;
static float area(int64_t id){const ort_bvh_node&n=B[id];float x=n.hi[0]-n.lo[0],y=n.hi[1]-n.lo[1],z=n.hi[2]-n.lo[2];float a=x*y+y*z+z*x;return std::isfinite(a)?a:0;}
// ---- strategy 0: greedy largest area
static void kids_greedy(int64_t src,int64_t*k,int&nk){nk=0;k[nk++]=B[src].a;k[nk++]=B[src].b;while(nk<4){int pick=-1;float best=-1;for(int i=0;i<nk;i++)if(B[k[i]].kind==1&&area(k[i])>best){best=area(k[i]);pick=i;}if(pick<0)break;int64_t o=k[pick];k[pick]=B[o].a;k[nk++]=B[o].b;}}
// ---- strategy 1: SAH-optimal collapse (DP): cost[n][i] = min cost of representing subtree n by at most i+1 roots
static std::vector<float> C; static std::vector<uint8_t> SPLIT; // C[n*4+i]; SPLIT[n*4+i] = #roots given to left (0 = reuse i-1)
static float cleaf(int64_t n){return area(n)*(float)B[n].b*2.0f;}
static void dp(){C.assign(NB*4,0);SPLIT.assign(NB*4,0);for(int64_t n=0;n<NB;n++){if(B[n].kind==0){for(int i=0;i<4;i++)C[n*4+i]=cleaf(n);continue;}int64_t l=B[n].a,r=B[n].b;
  // distribute j roots (2..4) among children
  float dist[5];uint8_t ds[5];for(int j=2;j<=4;j++){dist[j]=1e30f;for(int k=1;k<j;k++){float c=C[l*4+k-1]+C[r*4+j-k-1];if(c<dist[j]){dist[j]=c;ds[j]=(uint8_t)k;}}}
  C[n*4+0]=dist[4]+area(n)*1.0f; SPLIT[n*4+0]=ds[4];
  for(int i=1;i<4;i++){float cd=dist[i+1];if(cd<C[n*4+i-1]){C[n*4+i]=cd;SPLIT[n*4+i]=ds[i+1];}else{C[n*4+i]=C[n*4+i-1];SPLIT[n*4+i]=0;}}}}
// collect the roots representing subtree n with budget i+1
static void collect(int64_t n,int i,int64_t*k,int&nk){if(B[n].kind==0){k[nk++]=n;return;}if(i==0){k[nk++]=n;return;}uint8_t s=SPLIT[n*4+i];if(s==0){collect(n,i-1,k,nk);return;}collect(B[n].a,s-1,k,nk);collect(B[n].b,i+1-s-1,k,nk);}
static void kids_dp(int64_t src,int64_t*k,int&nk){nk=0;uint8_t s=SPLIT[src*4+0];collect(B[src].a,s-1,k,nk);collect(B[src].b,4-s-1,k,nk);}
static std::vector<WNode> build(int strat){std::vector<WNode> out;struct P{int64_t src;int dst;};std::vector<P> q;out.emplace_back();q.push_back({NB-1,0});
 for(size_t qi=0;qi<q.size();qi++){P cur=q[qi];int64_t k[8];int nk=0;if(B[cur.src].kind==0){k[nk++]=cur.src;}else if(strat==0)kids_greedy(cur.src,k,nk);else kids_dp(cur.src,k,nk);
  WNode w;for(int i=0;i<4;i++){for(int a=0;a<3;a++){w.lo[i][a]=INFINITY;w.hi[i][a]=-INFINITY;}w.child[i]=EMPTY;}
  for(int i=0;i<nk;i++){const ort_bvh_node&c=B[k[i]];memcpy(w.lo[i],c.lo,12);memcpy(w.hi[i],c.hi,12);if(c.kind==0){w.child[i]=~(int32_t)((c.a<<3)|c.b);}else{w.child[i]=(int)out.size();out.emplace_back();q.push_back({k[i],w.child[i]});}}
  out[cur.dst]=w;}return out;}
static const ort_triangle* T;
static bool tri_hit(const float*o,const float*d,const ort_triangle&t,double&tt){double e1[3]={t.u[0],t.u[1],t.u[2]},e2[3]={t.v[0],t.v[1],t.v[2]};double p[3]={d[1]*e2[2]-d[2]*e2[1],d[2]*e2[0]-d[0]*e2[2],d[0]*e2[1]-d[1]*e2[0]};double det=e1[0]*p[0]+e1[1]*p[1]+e1[2]*p[2];if(det==0)return false;double inv=1/det;double s[3]={o[0]-t.p[0],o[1]-t.p[1],o[2]-t.p[2]};double u=(s[0]*p[0]+s[1]*p[1]+s[2]*p[2])*inv;if(u<0||u>1)return false;double qv[3]={s[1]*e1[2]-s[2]*e1[1],s[2]*e1[0]-s[0]*e1[2],s[0]*e1[1]-s[1]*e1[0]};double v=(d[0]*qv[0]+d[1]*qv[1]+d[2]*qv[2])*inv;if(v<0||u+v>1)return false;tt=(e2[0]*qv[0]+e2[1]*qv[1]+e2[2]*qv[2])*inv;return tt>0;}
static void trace(const std::vector<WNode>&W,const ort_ray&r,long&visits,long&tris,long&leafs){float inv[3]={1/r.d[0],1/r.d[1],1/r.d[2]};double best=INFINITY;struct E{int n;float d;};E st[256];int sp=0;int cur=0;
 for(;;){if(cur>=0){visits++;const WNode&w=W[cur];float dd[4];int cc[4];int nh=0;for(int i=0;i<4;i++){float tn=0,tf=(float)best;bool ok=w.child[i]!=EMPTY;for(int a=0;a<3&&ok;a++){float t0=(w.lo[i][a]-r.o[a])*inv[a],t1=(w.hi[i][a]-r.o[a])*inv[a];if(t0>t1)std::swap(t0,t1);tn=std::max(tn,t0);tf=std::min(tf,t1);if(!(tn<=tf))ok=false;}if(ok){dd[nh]=tn;cc[nh]=w.child[i];nh++;}}
   for(int i=0;i<nh;i++)for(int j=i+1;j<nh;j++)if(dd[j]<dd[i]){std::swap(dd[i],dd[j]);std::swap(cc[i],cc[j]);}
   for(int i=nh-1;i>=1;i--)st[sp++]={cc[i],dd[i]};
   if(nh>0){cur=cc[0];if(cur<0){/*leaf*/}}else{cur=EMPTY;}
   if(nh>0&&cur>=0)continue;
   if(nh>0){uint32_t code=(uint32_t)~cur;uint32_t first=code>>3,cnt=code&7;leafs++;for(uint32_t i=0;i<cnt;i++){tris++;double t;if(tri_hit(r.o,r.d,T[first+i],t)&&t<best)best=t;}}
   cur=EMPTY;while(sp>0){E e=st[--sp];if(e.d<=best){cur=e.n;break;}}
   if(cur==EMPTY)return;
   if(cur<0){/* popped a leaf */ uint32_t code=(uint32_t)~cur;uint32_t first=code>>3,cnt=code&7;leafs++;for(uint32_t i=0;i<cnt;i++){tris++;double t;if(tri_hit(r.o,r.d,T[first+i],t)&&t<best)best=t;}
     // continue popping
     for(;;){cur=EMPTY;while(sp>0){E e=st[--sp];if(e.d<=best){cur=e.n;break;}}if(cur==EMPTY)return;if(cur>=0)break;uint32_t code2=(uint32_t)~cur;uint32_t f2=code2>>3,c2=code2&7;leafs++;for(uint32_t i=0;i<c2;i++){tris++;double t;if(tri_hit(r.o,r.d,T[f2+i],t)&&t<best)best=t;}}
   }
  } else return; }
}

// ---- 8-wide: greedy collapse, three traversal orders
struct W8 { float lo[8][3], hi[8][3]; int32_t child[8]; float cx[8][3]; };
static std::vector<W8> build8(){std::vector<W8> out;struct P{int64_t src;int dst;};std::vector<P> q;out.emplace_back();q.push_back({NB-1,0});
 for(size_t qi=0;qi<q.size();qi++){P cur=q[qi];int64_t k[8];int nk=0;if(B[cur.src].kind==0){k[nk++]=cur.src;}else{k[nk++]=B[cur.src].a;k[nk++]=B[cur.src].b;while(nk<8){int pick=-1;float best=-1;for(int i=0;i<nk;i++)if(B[k[i]].kind==1&&area(k[i])>best){best=area(k[i]);pick=i;}if(pick<0)break;int64_t o=k[pick];k[pick]=B[o].a;k[nk++]=B[o].b;}}
  W8 w;for(int i=0;i<8;i++){for(int a=0;a<3;a++){w.lo[i][a]=INFINITY;w.hi[i][a]=-INFINITY;}w.child[i]=EMPTY;}
  // octant slot assignment (greedy, as build_wide8_bvh)
  int slot_of[8],kid_at[8];for(int i=0;i<8;i++){slot_of[i]=-1;kid_at[i]=-1;}const ort_bvh_node&pn=B[cur.src];float pc[3];for(int a=0;a<3;a++)pc[a]=0.5f*(pn.lo[a]+pn.hi[a]);float cost[8][8];
  for(int i=0;i<nk;i++){const ort_bvh_node&c=B[k[i]];float d[3];for(int a=0;a<3;a++)d[a]=0.5f*(c.lo[a]+c.hi[a])-pc[a];for(int sl=0;sl<8;sl++)cost[i][sl]=((sl&1)?d[0]:-d[0])+((sl&2)?d[1]:-d[1])+((sl&4)?d[2]:-d[2]);}
  for(int round=0;round<nk;round++){int bi=-1,bs=-1;float bc=-INFINITY;for(int i=0;i<nk;i++){if(slot_of[i]>=0)continue;for(int sl=0;sl<8;sl++)if(kid_at[sl]<0&&cost[i][sl]>bc){bc=cost[i][sl];bi=i;bs=sl;}}slot_of[bi]=bs;kid_at[bs]=bi;}
  for(int sl=0;sl<8;sl++){if(kid_at[sl]<0)continue;const ort_bvh_node&c=B[k[kid_at[sl]]];memcpy(w.lo[sl],c.lo,12);memcpy(w.hi[sl],c.hi,12);if(c.kind==0){w.child[sl]=~(int32_t)((c.a<<3)|c.b);}else{w.child[sl]=(int)out.size();out.emplace_back();q.push_back({k[kid_at[sl]],w.child[sl]});}}
  out[cur.dst]=w;}return out;}
static inline void leaf_test(int32_t ref,const ort_ray&r,double&best,long&tris,long&leafs){uint32_t code=(uint32_t)~ref;uint32_t first=code>>3,cnt=code&7;leafs++;for(uint32_t i=0;i<cnt;i++){tris++;double t;if(tri_hit(r.o,r.d,T[first+i],t)&&t<best)best=t;}}
static inline bool box8(const W8&w,int i,const ort_ray&r,const float*inv,double best,float&tn){if(w.child[i]==EMPTY)return false;tn=0;float tf=(float)best;for(int a=0;a<3;a++){float t0=(w.lo[i][a]-r.o[a])*inv[a],t1=(w.hi[i][a]-r.o[a])*inv[a];if(t0>t1)std::swap(t0,t1);tn=std::max(tn,t0);tf=std::min(tf,t1);if(!(tn<=tf))return false;}return true;}
// mode 0: exact distance order + per-entry cull (leaves are entries too); mode 1: octant order, leaves of a node tested
// right after the visit (traverse8.cuh); mode 2: octant order, leaves ordered among the inner children
static void trace8(const std::vector<W8>&W,const ort_ray&r,int mode,long&visits,long&tris,long&leafs){float inv[3]={1/r.d[0],1/r.d[1],1/r.d[2]};double best=INFINITY;int oct=(r.d[0]<0?1:0)|(r.d[1]<0?2:0)|(r.d[2]<0?4:0);
 struct E{int n;float d;};std::vector<E> st;st.reserve(256);st.push_back({0,0.f});
 while(!st.empty()){E e=st.back();st.pop_back();if(mode==0&&e.d>best)continue;if(e.n<0){leaf_test(e.n,r,best,tris,leafs);continue;}
  visits++;const W8&w=W[e.n];E h[8];int nh=0;
  if(mode==0){for(int i=0;i<8;i++){float tn;if(box8(w,i,r,inv,best,tn))h[nh++]={w.child[i],tn};}std::sort(h,h+nh,[](const E&a,const E&b){return a.d<b.d;});for(int i=nh-1;i>=0;i--)st.push_back(h[i]);}
  else{ // octant order: ascending (slot ^ oct)
   for(int p=0;p<8;p++){int sl=p^oct;float tn;if(box8(w,sl,r,inv,best,tn)){if(mode==1&&w.child[sl]<0){leaf_test(w.child[sl],r,best,tris,leafs);}else h[nh++]={w.child[sl],tn};}}
   for(int i=nh-1;i>=0;i--)st.push_back(h[i]);}
 }}
int main(int argc,char**argv){std::string base=argv[1];auto tb=slurp((base+"_tris.bin").c_str());auto bb=slurp((base+"_bvh.bin").c_str());T=(const ort_triangle*)tb.data();B=(const ort_bvh_node*)bb.data();NB=bb.size()/sizeof(ort_bvh_node);dp();
 for(int strat=0;strat<2;strat++){auto W=build(strat);double sa=0;for(auto&w:W)for(int i=0;i<4;i++)if(w.child[i]>=0){float x=w.hi[i][0]-w.lo[i][0],y=w.hi[i][1]-w.lo[i][1],z=w.hi[i][2]-w.lo[i][2];sa+=x*y+y*z+z*x;}
  for(const char* rs:{"_primary.bin","_bounce.bin"}){auto rb=slurp((base+rs).c_str());const ort_ray*R=(const ort_ray*)rb.data();size_t n=rb.size()/sizeof(ort_ray);long v=0,t=0,l=0;for(size_t i=0;i<n;i++)trace(W,R[i],v,t,l);
   printf("%s strat=%s nodes=%zu innerSA=%.4g %s: visits/ray %.2f leaves/ray %.2f tris/ray %.2f\n",argv[1],strat?"dp":"greedy",W.size(),sa,rs,(double)v/n,(double)l/n,(double)t/n);}}
 auto W8v=build8();
 for(int mode=0;mode<3;mode++)for(const char* rs:{"_primary.bin","_bounce.bin"}){auto rb=slurp((base+rs).c_str());const ort_ray*R=(const ort_ray*)rb.data();size_t n=rb.size()/sizeof(ort_ray);long v=0,t=0,l=0;for(size_t i=0;i<n;i++)trace8(W8v,R[i],mode,v,t,l);
   printf("%s 8-wide nodes=%zu mode=%s %s: visits/ray %.2f leaves/ray %.2f tris/ray %.2f\n",argv[1],W8v.size(),mode==0?"exact-order+cull":mode==1?"octant,leaves-immediately":"octant,leaves-ordered",rs,(double)v/n,(double)l/n,(double)t/n);}
}
