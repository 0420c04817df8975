// Self-test of the process-per-rank MPI stand-in (tests/test_ref_grid.py builds and runs it on 1, 4 and 9 ranks).
#include <mpi.h>
#include <vector>
#include <cstdio>
#include <cassert>
static void user_max(void* in, void* io, int* n, MPI_Datatype*) { long long* a=(long long*)in; long long* b=(long long*)io; for(int i=0;i<*n;++i) b[i]=a[i]>b[i]?a[i]:b[i]; }
int main(int argc, char** argv) {
  MPI_Init(&argc,&argv); int r,p; MPI_Comm_rank(MPI_COMM_WORLD,&r); MPI_Comm_size(MPI_COMM_WORLD,&p);
  int q=1; while(q*q<p) ++q; assert(q*q==p);
  MPI_Comm row,col; MPI_Comm_split(MPI_COMM_WORLD, r/q, r%q, &row); MPI_Comm_split(MPI_COMM_WORLD, r%q, r/q, &col);
  int rr,cr; MPI_Comm_rank(row,&rr); MPI_Comm_rank(col,&cr); assert(rr==r%q && cr==r/q);
  // bcast along rows
  for (int root=0; root<q; ++root){ std::vector<double> v(1000, rr==root? r+0.5 : -1); MPI_Bcast(v.data(),1000,MPI_DOUBLE,root,row); assert(v[999]==(r/q)*q+root+0.5); }
  long long s=r+1, tot=0; MPI_Allreduce(&s,&tot,1,MPI_LONG_LONG,MPI_SUM,MPI_COMM_WORLD); assert(tot==(long long)p*(p+1)/2);
  double d=r; MPI_Allreduce(MPI_IN_PLACE,&d,1,MPI_DOUBLE,MPI_MAX,col); assert(d==(q-1)*q+r%q);
  std::vector<int> ag(p); int me=r*7; MPI_Allgather(&me,1,MPI_INT,ag.data(),1,MPI_INT,MPI_COMM_WORLD); for(int i=0;i<p;++i) assert(ag[i]==i*7);
  // alltoallv: rank r sends i+1 ints of value r*100+i to rank i
  std::vector<int> sc(p),sd(p),rc(p),rd(p); int tots=0; for(int i=0;i<p;++i){sc[i]=i+1; sd[i]=tots; tots+=i+1; rc[i]=r+1; rd[i]=i*(r+1);} 
  std::vector<int> sb(tots), rb(p*(r+1)); for(int i=0;i<p;++i) for(int k=0;k<=i;++k) sb[sd[i]+k]=r*100+i;
  MPI_Alltoallv(sb.data(),sc.data(),sd.data(),MPI_INT,rb.data(),rc.data(),rd.data(),MPI_INT,MPI_COMM_WORLD);
  for(int i=0;i<p;++i) for(int k=0;k<=r;++k) assert(rb[rd[i]+k]==i*100+r);
  // sendrecv ring + diag partner
  int nxt=(r+1)%p, prv=(r+p-1)%p, got=-1; MPI_Status st; MPI_Sendrecv(&r,1,MPI_INT,nxt,5,&got,1,MPI_INT,prv,5,MPI_COMM_WORLD,&st); assert(got==prv && st.MPI_SOURCE==prv);
  int cnt; MPI_Get_count(&st,MPI_INT,&cnt); assert(cnt==1);
  // irecv/isend out of order tags
  MPI_Request rq[2]; int a=-1,b=-1; MPI_Irecv(&a,1,MPI_INT,prv,2,MPI_COMM_WORLD,&rq[0]); MPI_Irecv(&b,1,MPI_INT,prv,1,MPI_COMM_WORLD,&rq[1]);
  int x1=r*10+1,x2=r*10+2; MPI_Request sq; MPI_Isend(&x1,1,MPI_INT,nxt,1,MPI_COMM_WORLD,&sq); MPI_Isend(&x2,1,MPI_INT,nxt,2,MPI_COMM_WORLD,&sq);
  MPI_Waitall(2,rq,MPI_STATUSES_IGNORE); assert(a==prv*10+2 && b==prv*10+1);
  // user op, scan, exscan, reduce_scatter, gather, scatter
  MPI_Op op; MPI_Op_create(user_max,1,&op); long long u=(r*37)%11, um=0; MPI_Allreduce(&u,&um,1,MPI_LONG_LONG,op,MPI_COMM_WORLD); long long ex=0; for(int i=0;i<p;++i) ex=std::max<long long>(ex,(i*37)%11); assert(um==ex);
  int one=1, sc2=0; MPI_Scan(&one,&sc2,1,MPI_INT,MPI_SUM,MPI_COMM_WORLD); assert(sc2==r+1); int ex2=-7; MPI_Exscan(&one,&ex2,1,MPI_INT,MPI_SUM,MPI_COMM_WORLD); assert(r==0||ex2==r);
  std::vector<int> cn(p,2), contrib(2*p); for(int i=0;i<2*p;++i) contrib[i]=r+i; int out2[2]; MPI_Reduce_scatter(contrib.data(),out2,cn.data(),MPI_INT,MPI_SUM,MPI_COMM_WORLD); assert(out2[0]==p*(p-1)/2+p*2*r);
  std::vector<int> g(p); MPI_Gather(&r,1,MPI_INT,g.data(),1,MPI_INT,0,MPI_COMM_WORLD); if(r==0) for(int i=0;i<p;++i) assert(g[i]==i);
  int sv=-1; MPI_Scatter(g.data(),1,MPI_INT,&sv,1,MPI_INT,0,MPI_COMM_WORLD); assert(sv==r);
  // comm_create of the diagonal, compare, dup
  MPI_Group wg,dg; MPI_Comm_group(MPI_COMM_WORLD,&wg); std::vector<int> diag(q); for(int i=0;i<q;++i) diag[i]=i*q+i; MPI_Group_incl(wg,q,diag.data(),&dg); MPI_Comm dc; MPI_Comm_create(MPI_COMM_WORLD,dg,&dc);
  if (r/q==r%q) { int ds; MPI_Comm_size(dc,&ds); assert(ds==q);} else assert(dc==MPI_COMM_NULL);
  MPI_Comm dup; MPI_Comm_dup(row,&dup); int cmp; MPI_Comm_compare(row,dup,&cmp); assert(cmp==MPI_CONGRUENT); MPI_Comm_compare(row,row,&cmp); assert(cmp==MPI_IDENT);
  if (q>1) { MPI_Comm_compare(row,col,&cmp); assert(cmp==MPI_UNEQUAL); }
  MPI_Barrier(MPI_COMM_WORLD); if(r==0) printf("cbmpi selftest ok on %d ranks\n",p);
  MPI_Finalize(); return 0; }
