// Parameter block of the aux-at-target MLP shared by head.cu and decode_grid.cu (HOST struct, DEVICE pointers).
#pragma once
#define CNP_MLP_MAX_LAYERS 6
struct cnp_mlp_params {
  const float* W[CNP_MLP_MAX_LAYERS];   // [out,in] row-major (torch Linear.weight)
  const float* b[CNP_MLP_MAX_LAYERS];   // [out]
  float* dW[CNP_MLP_MAX_LAYERS];        // (+=) gradients, backward only
  float* db[CNP_MLP_MAX_LAYERS];
  int dims[CNP_MLP_MAX_LAYERS + 1];     // dims[0] = Cf + Ca, dims[n_layers] = head inputs (2 | 4 | 5)
  int n_layers;
  int likelihood;                       // CNP_LIK_*: 0 heteroscedastic Gaussian, 1 Bernoulli-Gamma, 2 spikes-Beta
};
#define CNP_LIK_GAUSS 0
#define CNP_LIK_BERNOULLI_GAMMA 1
#define CNP_LIK_SPIKES_BETA 2
