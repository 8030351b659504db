#!/bin/sh
# Regenerates the golden traces in this directory by running the UNMODIFIED reference
# (oracle/_ref, built from /root/reference by oracle/Makefile) through
# tests/harness/scene_driver.cpp.  Traces are the reference's own outputs: callback pair
# sequence, contacts, joint feedback (lambda tap), body state and LCG seed per step.
set -e
cd "$(dirname "$0")/../.."
for p in single double; do
  d=oracle/_ref/driver_ref_$p
  $d --scene free6       --steps 20 --out tests/golden/free6_$p.trace
  $d --scene mixed       --steps 40 --out tests/golden/mixed_$p.trace
  $d --scene mixed_maxc4 --steps 40 --settle 80 --out tests/golden/mixed_maxc4_settle80_$p.trace
  $d --scene stack32     --steps 12 --settle 120 --worlds 2 --out tests/golden/stack32_w2_settle120_$p.trace
  $d --scene tower64     --steps 8 --settle 150 --out tests/golden/tower64_settle150_$p.trace
  $d --scene chain       --steps 30 --settle 60 --out tests/golden/chain_settle60_$p.trace
  $d --scene hinges      --steps 30 --settle 70 --out tests/golden/hinges_settle70_$p.trace
  $d --scene buggy       --steps 30 --settle 100 --out tests/golden/buggy_settle100_$p.trace
  $d --scene capsmix     --steps 20 --settle 90 --out tests/golden/capsmix_settle90_$p.trace
  $d --scene ragdoll     --steps 12 --settle 110 --out tests/golden/ragdoll_settle110_$p.trace
  $d --scene block64@sap --steps 6 --settle 20 --out tests/golden/block64_sap_settle20_$p.trace
  $d --scene mixed@sapz  --steps 30 --settle 60 --out tests/golden/mixed_sapz_settle60_$p.trace
  $d --scene mixed@simple --steps 30 --settle 60 --out tests/golden/mixed_simple_settle60_$p.trace
  $d --scene terrain_spheres --steps 25 --settle 70 --out tests/golden/terrain_spheres_settle70_$p.trace
  $d --scene terrain_boxes --steps 15 --settle 70 --out tests/golden/terrain_boxes_settle70_$p.trace
  $d --scene buggy_terrain --steps 30 --settle 90 --worlds 2 --out tests/golden/buggy_terrain_w2_settle90_$p.trace
  $d --scene terrain_capsules --steps 20 --settle 70 --out tests/golden/terrain_capsules_settle70_$p.trace
  $d --scene terrain_plane --steps 20 --settle 60 --out tests/golden/terrain_plane_settle60_$p.trace
  $d --scene terrain_capsules_pre --steps 20 --settle 70 --out tests/golden/terrain_capsules_pre_settle70_$p.trace
  $d --scene sliders --steps 30 --settle 60 --out tests/golden/sliders_settle60_$p.trace
  $d --scene universals --steps 30 --settle 60 --out tests/golden/universals_settle60_$p.trace
  $d --scene motors --steps 30 --settle 60 --out tests/golden/motors_settle60_$p.trace
  $d --scene pistons --steps 30 --settle 60 --out tests/golden/pistons_settle60_$p.trace
  $d --scene pus --steps 30 --settle 60 --out tests/golden/pus_settle60_$p.trace
  $d --scene cylmix --steps 20 --settle 90 --out tests/golden/cylmix_settle90_$p.trace
  $d --scene kinematic --steps 30 --settle 60 --out tests/golden/kinematic_settle60_$p.trace
  $d --scene nulljoint --steps 30 --settle 60 --out tests/golden/nulljoint_settle60_$p.trace
  $d --scene transforms --steps 25 --settle 80 --out tests/golden/transforms_settle80_$p.trace
  $d --scene bodyflags --steps 30 --settle 50 --out tests/golden/bodyflags_settle50_$p.trace
  $d --scene autodisable --steps 30 --settle 150 --out tests/golden/autodisable_settle150_$p.trace
  $d --scene contactmodes --steps 30 --settle 60 --out tests/golden/contactmodes_settle60_$p.trace
  $d --scene crashwall --steps 30 --settle 45 --out tests/golden/crashwall_settle45_$p.trace
  $d --scene autodisable_avg --steps 30 --settle 150 --out tests/golden/autodisable_avg_settle150_$p.trace
  # drop-in (callback) path only: ray colliders + capsule-trimesh, FDir1, per-call max-contacts
  $d --scene raycast --steps 20 --settle 40 --out tests/golden/raycast_settle40_$p.trace
  $d --scene raycast2 --steps 20 --settle 40 --out tests/golden/raycast2_settle40_$p.trace
  $d --scene raycyl --steps 20 --settle 40 --out tests/golden/raycyl_settle40_$p.trace
  $d --scene contactmodes_fdir1 --steps 25 --settle 60 --out tests/golden/contactmodes_fdir1_settle60_$p.trace
  $d --scene mixed_varmaxc --steps 30 --settle 40 --out tests/golden/mixed_varmaxc_settle40_$p.trace
  $d --scene nested --steps 25 --settle 60 --out tests/golden/nested_settle60_$p.trace
  $d --scene nested_dcollide --steps 25 --settle 60 --out tests/golden/nested_dcollide_settle60_$p.trace
  # large-world path (config 5): reference trace used in lock-step (--resync) by tests/test_large_world.py
  $d --scene pile_5x5x8 --steps 10 --settle 60 --out tests/golden/pile_5x5x8_large_settle60_$p.trace
done
# dWorldExportDIF text dumps written by the reference (tests/test_export_dif.py)
oracle/_ref/driver_ref_single --scene pistons --steps 25 --settle 20 --export-dif tests/golden/pistons_single.dif > /dev/null
oracle/_ref/driver_ref_double --scene motors --steps 25 --settle 20 --export-dif tests/golden/motors_double.dif > /dev/null
oracle/_ref/driver_ref_single --scene cylmix --steps 25 --settle 20 --export-dif tests/golden/cylmix_single.dif > /dev/null
oracle/_ref/driver_ref_double --scene cylmix --steps 25 --settle 20 --export-dif tests/golden/cylmix_double.dif > /dev/null
oracle/_ref/driver_ref_single --scene transforms --steps 25 --settle 20 --export-dif tests/golden/transforms_single.dif > /dev/null
# accessor probe (tests/test_api_probe.py): output of the probe linked against the reference
oracle/_ref/api_probe_ref_single > tests/golden/api_probe_single.txt
oracle/_ref/api_probe_ref_double > tests/golden/api_probe_double.txt
ls -la tests/golden
