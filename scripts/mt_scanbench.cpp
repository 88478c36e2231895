#include "../phet_b200/csrc/mt_replay.cpp"
#include <chrono>
#include <cstdio>
namespace phet { void set_error(const std::string& m){ fprintf(stderr,"ERR %s\n", m.c_str()); } }
template <typename E> void run(const char* name){
    OutRing r; r.ebytes=sizeof(E); r.region_elems=64*624; r.nreg=256; r.cap=r.nreg*r.region_elems; // 10.2M elems
    r.w = aligned_alloc(64,(r.cap+256)*sizeof(E)); r.ready.reset(new std::atomic<uint64_t>[r.nreg]);
    uint32_t key[624]; for(int i=0;i<624;i++) key[i]=i*2654435761u+1;
    phet_mt* e=mt_alloc(); e->load(key,0);
    E* w=(E*)r.w; for (size_t i=0;i<r.cap+128;i++) w[i]=(E)e->next32();
    for (size_t i=0;i<r.nreg;i++) r.ready[i].store(i+1);
    for (int rep=0;rep<3;rep++){
      RingSrc<E> src(&r,0);
      auto t0=std::chrono::steady_clock::now(); int perms=0;
      while (src.abs + 60000 < r.cap){ if (sizeof(E)==2 && getenv("FAST16")) count_draws16_avx512(src,25000); else shuffle_draws_from<uint16_t>(src,true,25000,(uint16_t*)nullptr); perms++; }
      double dt=std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count();
      printf("%s: %d perms, %.2f us/perm, %.3f ns/raw\n", name, perms, dt/perms*1e6, dt/src.abs*1e9);
    }
}
int main(){ run<uint16_t>("u16"); run<uint32_t>("u32"); }
