// Empty stand-in: THC/THC.h was removed from torch >= 1.11 but the reference still includes it
// (lib/pointgroup_ops/src/bfs_cluster/bfs_cluster.h:11, knn/knn.h:5, six pointops2 .cpp files) without using it.
#pragma once
