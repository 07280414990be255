// host-side kernel-launch counter behind odise_launch_count() (bench.py "gpu_launches")
#pragma once
namespace ob {
void count_launch(int n);
}
