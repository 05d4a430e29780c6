# Drop-in for the body of the loop in R/main.R:79-88 of the reference: one device call
# instead of maxiter x (5-8) .Call round trips.  Everything else in CausalStump()
# (check_inputs, norm_variables, kernel/optimizer construction, plots, return value)
# stays as in the reference.  In a forked worker (parallel::makeCluster(type="FORK"),
# test/sec5-3_Hill2011.R:611-616) CUDA is initialised lazily after the fork by the first
# b200_* call; pick the device with Sys.setenv(CUDA_VISIBLE_DEVICES = worker_rank %% n_gpu).
b200_optimise <- function(myKernel, myOptimizer, y, X, z, maxiter, tol, prior) {
  optim_code <- switch(class(myOptimizer)[1], AdamOpt = 0L, NadamOpt = 1L, NesterovOpt = 2L)
  lr <- myOptimizer$lr
  b1 <- if (optim_code == 2L) 0 else myOptimizer$beta1
  b2 <- if (optim_code == 2L) 0 else myOptimizer$beta2
  mom <- if (optim_code == 2L) myOptimizer$momentum else 0
  h <- b200_fit_create(as.integer(prior), optim_code, lr, b1, b2, mom, y, as.matrix(X), z, myKernel$w,
                       myKernel$parameters)
  res <- b200_fit_run(h, as.integer(maxiter), tol, myKernel$parameters)
  myKernel$parameters <- res$parameters
  attr(myKernel, "b200_handle") <- h   # predict()/treatment() reuse the device-resident K^-1
  list(iter = res$iter, stats = res$stats)
}
