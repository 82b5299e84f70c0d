## aptb200.R — R shim that keeps nimbleAPT's user surface on top of libaptb200.so.
##
## NOT exercised in this image (no R, no nimble: SURVEY.md F7).  It mirrors, name for name, what a user of the
## reference touches (nimbleAPT/NAMESPACE:3-9; vignette :127-141):
##   conf$addSampler(target, type = "sampler_RW_tempered" | "sampler_RW_block_tempered" | "sampler_slice_tempered" |
##                   "sampler_RW_multinomial_tempered", control = list(...))
##   buildAPT(conf, Temps, monitorTmax = TRUE, ULT = 1E6, thinPrintTemps = 1, ...)      APT_functions.R:242-248
##   cAPT$run(niter, reset, resetMV, resetTempering, simulateAll, time, adaptTemps, printTemps, tuneTemper1,
##            tuneTemper2, progressBar, thin, thin2)                                      APT_functions.R:336-348
##   as.matrix(cAPT$mvSamples), cAPT$logProbs, cAPT$tempTraj, cAPT$accCountSwap, cAPT$Temps
##   cAPT$calculateWAIC(nburnin)                                                         APT_functions.R:593-635
##   runAPT(cAPT, niter, thin, ...)   (named by the north star; not present in the reference, SURVEY.md F5)
## The sampler names are *recording* samplers: they capture target + control and never run in R.

.aptb200 <- new.env()

sampler_RW_tempered       <- structure(list(name = "sampler_RW_tempered"),       class = "aptb200_sampler_type")
sampler_RW_block_tempered <- structure(list(name = "sampler_RW_block_tempered"), class = "aptb200_sampler_type")
sampler_slice_tempered    <- structure(list(name = "sampler_slice_tempered"),    class = "aptb200_sampler_type")
sampler_RW_multinomial_tempered <- structure(list(name = "sampler_RW_multinomial_tempered"), class = "aptb200_sampler_type")

## Works with a nimble MCMCconf (conf$samplerConfs[[i]]$target / $name / $control, conf$monitors, conf$thin, conf$model)
## or with a plain list(model = list(code =, constants =, data =, inits =), samplers = list(...), monitors =, thin =).
buildAPT <- function(conf, Temps, monitorTmax = TRUE, ULT = 1E6, thinPrintTemps = 1, nLadders = 1, seed = 11L, device = 0L, functions = list(), ...) {
    if (missing(ULT)) message(paste("ULT set to", ULT))                                   # APT:254-257
    if (inherits(conf, "modelBaseClass")) conf <- nimble::configureMCMC(conf, ...)        # APT:249-250
    is_nimble <- inherits(conf, "MCMCconf")
    if (!is_nimble && !is.list(conf))
        stop("conf must either be a nimbleModel or a MCMCconf object (created by configureMCMC(...) )")  # APT:252
    if (is_nimble) {
        model     <- conf$model
        code      <- paste(deparse(model$getCode()), collapse = "\n")
        constants <- model$getConstants()
        dataNames <- unique(gsub("\\[.*", "", model$getNodeNames(dataOnly = TRUE)))
        data      <- lapply(setNames(dataNames, dataNames), function(v) model[[v]])
        parNames  <- setdiff(unique(gsub("\\[.*", "", model$getNodeNames(stochOnly = TRUE, includeData = FALSE))), dataNames)
        inits     <- lapply(setNames(parNames, parNames), function(v) model[[v]])
        samplers  <- lapply(conf$samplerConfs, function(s) list(target = paste0("c(", paste0("'", s$target, "'", collapse = ","), ")"),
                                                                  type = paste0("sampler_", sub("^sampler_", "", s$name)), control = s$control))
        monitors <- conf$monitors; monitors2 <- conf$monitors2; thin <- conf$thin; thin2 <- conf$thin2
    } else {
        code <- conf$model$code; constants <- conf$model$constants; data <- conf$model$data; inits <- conf$model$inits
        samplers <- conf$samplers; monitors <- conf$monitors; monitors2 <- conf$monitors2
        thin <- if (is.null(conf$thin)) 1L else conf$thin; thin2 <- if (is.null(conf$thin2)) 1L else conf$thin2
    }
    m <- .Call("aptR_model_create", code)
    ## user-supplied nimbleFunctions the model code calls (demo/APT_demo.R:42-71): `functions = list(k2theta = k2theta, ...)`;
    ## the text of their run functions is inlined by the library instead of being compiled by nimble
    for (nm in names(functions)) {
        f <- functions[[nm]]
        run <- if (is.function(f)) environment(f)$nfMethodRCobject$run else f   # [NIMBLE-UNVERIFIED] where nimble keeps `run`
        .Call("aptR_model_add_function", m, nm, if (is.character(run)) run else paste(deparse(run), collapse = "\n"))
    }
    for (nm in names(constants)) .Call("aptR_model_set_array", m, nm, 0L, constants[[nm]])
    for (nm in names(data))      .Call("aptR_model_set_array", m, nm, 1L, data[[nm]])
    for (nm in names(inits))     .Call("aptR_model_set_array", m, nm, 2L, inits[[nm]])
    for (s in samplers) {
        if (grepl("multinomial", s$type)) {
            if (is.null(s$control$useTempering)) stop("control$useTempering unspecified in sampler_RW_multinomial_tempered")  # APT:1015
        } else if (is.null(s$control$temperPriors) && !grepl("slice", s$type))           # APT:660, APT:781 (not read by the slice sampler, APT:909-912)
            stop(paste0("control$temperPriors unspecified in ", s$type))
        .Call("aptR_model_add_sampler", m, s$target, s$type, s$control)
    }
    .Call("aptR_model_set_monitors", m, 1L, as.character(monitors),  as.integer(thin), as.integer(thin2))
    .Call("aptR_model_set_monitors", m, 2L, as.character(monitors2), as.integer(thin), as.integer(thin2))
    Temps <- as.numeric(sort(Temps))                                                     # APT:266
    dots <- list(...)
    enableWAIC <- isTRUE(dots$enableWAIC) || isTRUE(conf$enableWAIC)                      # APT:259,329
    h <- .Call("aptR_build", m, Temps, monitorTmax, ULT, as.integer(nLadders), as.integer(seed), as.integer(device), enableWAIC)
    self <- new.env()
    self$handle <- h; self$model <- m; self$nTemps <- length(Temps); self$thinPrintTemps <- thinPrintTemps
    self$monitorNames  <- strsplit(.Call("aptR_model_names", m, 1L), "\n")[[1]]
    self$monitorNames2 <- strsplit(.Call("aptR_model_names", m, 2L), "\n")[[1]]
    info <- function(what) .Call("aptR_info", h, as.integer(what))
    get  <- function(what, rows, cols, ladder = 0L) .Call("aptR_get", h, as.integer(what), as.integer(ladder), rows, cols)
    ## samplesOut (extension): a numeric vector of nLadders * (niter %/% thin) * length(monitorNames) doubles that receives this
    ## run's rows of mvSamples while the run is still sampling; ladder l is t(matrix(block l, length(monitorNames)))
    self$run <- function(niter, reset = TRUE, resetMV = FALSE, resetTempering = FALSE, simulateAll = FALSE, time = FALSE,
                         adaptTemps = TRUE, printTemps = FALSE, tuneTemper1 = 10, tuneTemper2 = 1, progressBar = TRUE,
                         thin = -1L, thin2 = -1L, samplesOut = NULL) {
        .Call("aptR_run", h, niter, c(reset, resetMV, resetTempering, simulateAll, time, adaptTemps, as.logical(printTemps)),
              c(tuneTemper1, tuneTemper2), as.integer(c(thin, thin2)), samplesOut)
        invisible(NULL)
    }
    ## extensions of the C ABI that have no counterpart in the reference's object: new initial values / data for the same
    ## compiled model (the reference would assign into the compiled model, cModel$beta <- ...), and model$calculate() at
    ## given states (rows of theta), evaluated on the device
    self$setInits  <- function(theta, ladder = -1L) invisible(.Call("aptR_set_inits", h, as.numeric(theta), as.integer(ladder)))
    self$setData   <- function(name, values) invisible(.Call("aptR_set_array", h, name, values))
    self$calculate <- function(theta) { lp <- .Call("aptR_logprob", h, t(rbind(theta))); list(total = lp[1, ], groups = t(lp[-1, , drop = FALSE])) }
    self$calculateWAIC <- function(nburnin = 0, ladder = 0L) {                            # APT:593-635
        if (!enableWAIC) {
            print("Error: must set enableWAIC = TRUE in buildAPT to use the method calcualteWAIC. See help(buildAPT) and help(buildMCMC) for additional information.")
            return(NaN)
        }
        w <- .Call("aptR_waic", h, nburnin, as.integer(ladder))
        if (w[1] == -Inf) print("Error: need more than one post burn-in MCMC samples")
        else if (is.nan(w[1])) print("WAIC was calculated as NaN.  You may need to add monitors to model latent states, in order for a valid WAIC calculation.")
        w[1]
    }
    mk <- function(what, names, rowsInfo) structure(list(fetch = function(ladder = 0L) {
        x <- get(what, info(rowsInfo), length(names), ladder); colnames(x) <- names; x }), class = "aptb200_samples")
    self$mvSamples      <- mk(0L, self$monitorNames, 7L)
    self$mvSamples2     <- mk(1L, self$monitorNames2, 8L)
    self$mvSamplesTmax  <- mk(2L, self$monitorNames, 7L)
    self$mvSamples2Tmax <- mk(3L, self$monitorNames2, 8L)
    makeActiveBinding("logProbs",     function() drop(get(4L, info(7L), 1)), self)
    makeActiveBinding("tempTraj",     function() get(5L, info(9L), self$nTemps), self)
    makeActiveBinding("Temps",        function() drop(get(6L, 1, self$nTemps)), self)
    makeActiveBinding("accCountSwap", function() get(7L, self$nTemps, self$nTemps), self)
    class(self) <- "aptb200"
    self
}

as.matrix.aptb200_samples <- function(x, ...) x$fetch()
compileNimble_aptb200 <- function(x, ...) x     # buildAPT already lowered the model and compiled it with nvcc
runAPT <- function(cAPT, niter, thin = -1L, ...) { cAPT$run(niter, thin = thin, ...); as.matrix(cAPT$mvSamples) }
plotTempTraj <- function(cAPT) {                                                        # APT:1290-1308
    traj <- cAPT$tempTraj
    plot(c(1, nrow(traj)), range(traj), typ = "n", xlab = "Iteration", ylab = "Temperature", log = "y")
    for (ii in seq_len(ncol(traj))) lines(traj[, ii], col = rainbow(ncol(traj))[ii])
}
