/*
 * ref_births.hpp — TEST INFRASTRUCTURE, NOT PRODUCT.
 * Drives the reference's OWN RBPHDFilter::addBirthGaussians() (include/RBPHDFilter.hpp:1000-1080) on injected state
 * (included by ref_harness.cpp / ref_harness_vp.cpp after the reference headers, with private members opened).
 */
#ifndef REF_BIRTHS_HPP
#define REF_BIRTHS_HPP

template <class F, int D>
int ref_birth_candidates_run(F* f, phd_birth_io* io, const typename F::TMeasurement::Mat& R) {
  constexpr int NC = D * (D + 1) / 2;
  typedef typename F::TLandmark TLandmark;
  typedef typename F::TMeasurement TMeasurement;
  typedef typename F::TPose TPose;
  const int N = io->N;
  f->config.birthGaussianWeight_ = 0.25;   /* any weight: the callers take it from their own configuration */
  f->config.birthGaussianMeasurementCountThreshold_ = io->count_thr;
  f->config.birthGaussianMeasurementCheckThreshold_ = io->check_thr;
  f->config.birthGaussianCurrentMeasurementCountThreshold_ = io->cur_count_thr;
  f->config.birthGaussianMeasurementSupportDist_ = io->support_dist;
  f->resampleOccured_ = io->resample_occurred != 0;
  std::vector<TMeasurement> Z;
  for (int z = 0; z < io->nZ; z++) {
    typename TMeasurement::Vec zv;
    for (int d = 0; d < D; d++) zv(d) = io->Z[(size_t)D * z + d];
    Z.push_back(TMeasurement(zv, R));
  }
  f->setMeasurements(Z);
  for (int i = 0; i < N; i++) {
    typename TPose::Vec x;
    x << io->pose[3 * i], io->pose[3 * i + 1], io->pose[3 * i + 2];
    typename TPose::Mat Sx;
    Sx.setZero();
    if (D == 2 && io->pose_cov) {
      const double* s = io->pose_cov + 6 * (size_t)i;
      Sx << s[0], s[1], s[2], s[1], s[3], s[4], s[2], s[4], s[5];
    }
    TPose p(x, Sx);
    f->setParticlePose(i, p);
    f->particleSet_[i]->setId(i);
    f->particleSet_[i]->setParentId(io->parent ? (unsigned)io->parent[i] : (unsigned)i);
    f->unused_measurements_[i].clear();
    for (int z = 0; z < io->nZ; z++)
      if ((io->unused[i] >> z) & 1ull) f->unused_measurements_[i].push_back(z);
    f->nLandmarksInFOV_[i] = (unsigned)io->nfov[i];
    f->birthGaussians_[i].clear();
    for (int k = 0; k < io->cand_n[i]; k++) {
      const size_t s = (size_t)i * io->cand_cap + k;
      typename TLandmark::Vec lx;
      typename TLandmark::Mat lS;
      for (int d = 0; d < D; d++) lx(d) = io->cand_mean[s * D + d];
      for (int r = 0, q = 0; r < D; r++)
        for (int c = r; c < D; c++, q++) lS(r, c) = lS(c, r) = io->cand_cov[s * NC + q];
      typename F::BirthGaussianCandidate c;
      c.set(lx, lS);
      c.nSupportingMeasurements = (unsigned)io->cand_support[s];
      c.nChecks = (unsigned)io->cand_checks[s];
      f->birthGaussians_[i].push_back(c);
    }
  }
  f->addBirthGaussians();
  int rc = 0;
  for (int i = 0; i < N; i++) {
    io->unused[i] = 0;
    for (size_t u = 0; u < f->unused_measurements_[i].size(); u++) io->unused[i] |= 1ull << f->unused_measurements_[i][u];
    typename F::TGM* gm = f->getParticle(i)->getData().get();
    int n = 0;
    for (size_t m = 0; m < gm->gList_.size(); m++) {
      if (gm->gList_[m].landmark == NULL) continue;
      if (n < io->add_cap) {
        typename TLandmark::Vec lx;
        typename TLandmark::Mat lS;
        gm->gList_[m].landmark->get(lx, lS);
        const size_t s = (size_t)i * io->add_cap + n;
        for (int d = 0; d < D; d++) io->add_mean[s * D + d] = lx(d);
        for (int r = 0, q = 0; r < D; r++)
          for (int c = r; c < D; c++, q++) io->add_cov[s * NC + q] = lS(r, c);
      }
      n++;
    }
    io->add_n[i] = n;
    int k = 0;
    for (typename std::list<typename F::BirthGaussianCandidate>::iterator it = f->birthGaussians_[i].begin();
         it != f->birthGaussians_[i].end(); ++it, ++k) {
      if (k >= io->cand_cap) { rc = -4; break; }
      const size_t s = (size_t)i * io->cand_cap + k;
      typename TLandmark::Vec lx;
      typename TLandmark::Mat lS;
      it->get(lx, lS);
      for (int d = 0; d < D; d++) io->cand_mean[s * D + d] = lx(d);
      for (int r = 0, q = 0; r < D; r++)
        for (int c = r; c < D; c++, q++) io->cand_cov[s * NC + q] = lS(r, c);
      io->cand_support[s] = (int32_t)it->nSupportingMeasurements;
      io->cand_checks[s] = (int32_t)it->nChecks;
    }
    io->cand_n[i] = (int32_t)f->birthGaussians_[i].size();
  }
  return rc;
}
#endif
