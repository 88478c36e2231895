#include "../phet_b200/csrc/mt_replay.cpp"
#include <chrono>
#include <cstdio>
namespace phet { void set_error(const std::string& m){ fprintf(stderr,"ERR %s\n", m.c_str()); } }
int main(int argc, char** argv){
    int workers = argc>1 ? atoi(argv[1]) : 4;
    uint32_t key[624]; for(int i=0;i<624;i++) key[i]=i*2654435761u+1;
    phet_mt* e; phet_mt_create(key,624,&e);
    const int n=25000, per=160, nblk=24, slots=4;
    std::vector<std::vector<uint16_t>> buf(slots, std::vector<uint16_t>((size_t)2*per*n, 1));
    for (int rep=0; rep<2; rep++){
      auto t0=std::chrono::steady_clock::now();
      phet_mt_par* par; phet_mt_par_begin(e, workers, n, &par);
      std::vector<int64_t> tk(nblk);
      for (int b=0;b<nblk;b++){
          if (b>=slots) phet_mt_par_wait(par, tk[b-slots]);
          uint16_t* d=getenv("SKIP")?nullptr:buf[b%slots].data();
          phet_mt_job j[2]={{n,d,(int64_t)n*2,2},{n,d?d+(size_t)per*n:nullptr,(int64_t)n*2,2}};
          phet_mt_par_submit(par,j,2,per,&tk[b]);
      }
      for (int b=nblk-slots;b<nblk;b++) phet_mt_par_wait(par, tk[b]);
      phet_mt_par_end(par);
      double dt=std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count();
      printf("workers=%d: %.2f us/perm -> cfg4 %.3f s\n", workers, dt/(2.0*per*nblk)*1e6, dt/(2.0*per*nblk)*62000);
    }
    if (argc>2) return 0;   // parallel part only
    auto t0=std::chrono::steady_clock::now();
    for (int b=0;b<nblk;b++){ uint16_t* d=buf[b%slots].data(); phet_mt_job j[2]={{n,d,(int64_t)n*2,2},{n,d+(size_t)per*n,(int64_t)n*2,2}}; phet_mt_shuffle_rounds(e,j,2,per);} 
    double dt=std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count();
    printf("sequential store: %.2f us/perm -> cfg4 %.3f s\n", dt/(2.0*per*nblk)*1e6, dt/(2.0*per*nblk)*62000);
    t0=std::chrono::steady_clock::now();
    for (int b=0;b<nblk;b++){ phet_mt_job j[2]={{n,nullptr,0,2},{n,nullptr,0,2}}; phet_mt_shuffle_rounds(e,j,2,per);} 
    dt=std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count();
    printf("sequential count-only: %.2f us/perm -> cfg4 %.3f s\n", dt/(2.0*per*nblk)*1e6, dt/(2.0*per*nblk)*62000);
}
