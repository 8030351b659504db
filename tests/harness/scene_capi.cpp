// scene_capi — C entry point that builds BASELINE.json scenes (scenes.h) through the
// ODE C API and binds them into one batch, so bench.py / tests can drive the product
// library from Python with a handful of ctypes calls instead of ~10^6.
// Links against libode_b200_<prec>.so (or the hostsim test library).  Not product code.
#include <stdio.h>
#include "scenes.h"

static std::vector<std::vector<SceneWorld> *> g_keep;

extern "C" dBatchID ob_scene_build_batch(const char *scene, int nworlds, int world0, int contacts_cap, int device) {
  std::vector<SceneWorld> *worlds = new std::vector<SceneWorld>(nworlds);
  ScenePolicy pol;
  for (int w = 0; w < nworlds; w++)
    if (scene_build(scene, (*worlds)[w], world0 + w, pol)) { delete worlds; return 0; }
  std::vector<dWorldID> wv(nworlds);
  std::vector<dSpaceID> sv(nworlds);
  std::vector<uint32_t> seeds(nworlds);
  for (int w = 0; w < nworlds; w++) { wv[w] = (*worlds)[w].world; sv[w] = (*worlds)[w].space; seeds[w] = (*worlds)[w].seed; }
  dBatchDesc desc;
  memset(&desc, 0, sizeof(desc));
  desc.max_contacts_per_world = contacts_cap;
  desc.device = device;
  dBatchID B = dBatchCreate(nworlds, wv.data(), sv.data(), &desc);
  if (!B) return 0;
  dBatchContactPolicy bp[2];
  memset(bp, 0, sizeof(bp));
  int nrows = 0;
  if (pol.sphere_mu > 0) {   // pairs with a sphere first (category bit SCENE_CAT_SPHERE), then the catch-all row
    bp[0].cat_mask1 = SCENE_CAT_SPHERE; bp[0].cat_mask2 = ~0ul;
    bp[0].max_contacts = pol.max_contacts; bp[0].skip_if_connected = pol.skip_if_connected;
    bp[0].surface = pol.surface; bp[0].surface.mu = pol.sphere_mu;
    nrows = 1;
  }
  bp[nrows].cat_mask1 = bp[nrows].cat_mask2 = ~0ul;
  bp[nrows].max_contacts = pol.max_contacts;
  bp[nrows].skip_if_connected = pol.skip_if_connected;
  bp[nrows].surface = pol.surface;
  nrows++;
  dBatchSetContactPolicy(B, bp, nrows);
  dBatchSetSeeds(B, seeds.data());
  g_keep.push_back(worlds);
  return B;
}
