/*
 * Plain-C client of the sylph_b200 C ABI (include/sylph_b200.h): what a non-Python host binds.
 *
 *   gcc -std=c99 -Wall -Wextra -pedantic -Iinclude examples/c_client.c \
 *       -Lsylph_few_shot_detection_b200 -lsylph_b200 -Wl,-rpath,$PWD/sylph_few_shot_detection_b200 -o build/c_client
 *
 * Without an sm_100 device sylph_create fails (there is no CPU fallback) and the program reports that and exits 0;
 * with one it creates a context for the COCO Meta-FCOS configuration, shows the error path of an entry point that is
 * called too early (weights not loaded), and sets up / tears down a single-rank class-code exchange.
 */
#include <stdio.h>
#include <string.h>

#include "sylph_b200.h"

int main(void) {
    sylph_model_config cfg;
    sylph_ctx* ctx = NULL;
    int rc;
    memset(&cfg, 0, sizeof(cfg));
    cfg.resnet_depth = 50;
    cfg.num_cls_convs = 4;
    cfg.num_box_convs = 4;
    cfg.use_scale = 1;
    cfg.box_quality = 1;
    cfg.pre_nms_topk = 1000;
    cfg.post_nms_topk = 100;
    cfg.inference_thresh = 0.05f;
    cfg.nms_thresh = 0.6f;
    cfg.prior_prob = 0.01f;
    cfg.pixel_mean[0] = 103.53f; cfg.pixel_mean[1] = 116.28f; cfg.pixel_mean[2] = 123.675f;
    cfg.pixel_std[0] = cfg.pixel_std[1] = cfg.pixel_std[2] = 1.0f;
    cfg.cg_tower_layers = 2;
    cfg.cg_post_norm = 1;
    cfg.cg_conv_l2_norm = 1;
    cfg.cg_bias_layer = 1;
    cfg.cg_use_bias = 1;
    cfg.cg_has_conv_scale = 1;

    printf("library: %s\n", sylph_version());
    printf("code row: %d floats, detection row: %d floats, IPC handle: %d bytes\n", SYLPH_CODE_STRIDE, SYLPH_DET_STRIDE,
           SYLPH_IPC_HANDLE_BYTES);
    rc = sylph_create(&ctx, 0, &cfg);
    if (rc != 0 || ctx == NULL) {
        printf("sylph_create: status %d -- no sm_100 device here, and there is no CPU fallback\n", rc);
        return 0;
    }
    /* an entry point called before the weights are loaded reports through the status + message convention */
    rc = sylph_normalize_codes(ctx, NULL, NULL, 1, NULL);
    printf("sylph_normalize_codes before sylph_finalize_weights: status %d (%s)\n", rc, sylph_last_error(ctx));
    {
        uint8_t handle[SYLPH_IPC_HANDLE_BYTES];
        int timed_out = -1;
        int64_t rows = -1;
        rc = sylph_exchange_create(ctx, 1, 0, 64, handle);
        if (rc == 0) rc = sylph_exchange_connect(ctx, NULL);
        if (rc == 0) rc = sylph_exchange_status(ctx, &timed_out, &rows);
        printf("single-rank exchange: status %d, timed_out %d, rows %ld\n", rc, timed_out, (long)rows);
        sylph_exchange_destroy(ctx);
    }
    sylph_destroy(ctx);
    return 0;
}
