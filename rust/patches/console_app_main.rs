//! console_app/src/main.rs with the `--backend cuda` switch.  Lines marked `// +` are new; everything else is
//! the reference's file (main.rs:1-95) and is meant to stay byte for byte, in particular the tonemap / PNG block
//! (main.rs:66-94), which consumes the same `Vec<Pixel>` from either back end.
//! NOT COMPILED in the repository that ships it (no Rust toolchain there); the C++ mirror
//! raytracer-weekend_b200/host/console_app.cpp runs the same flow.
mod scenes;

use clap::Parser;
use image::{Rgb, RgbImage};
use indicatif::{ParallelProgressIterator, ProgressBar, ProgressIterator, ProgressStyle};
use rand::thread_rng;
use rayon::prelude::*;
use raytracer_weekend_lib::{Pixel, Raytracer};
use scenes::Scene;

const CRATE_VERSION: &str = env!("CARGO_PKG_VERSION");
const CRATE_AUTHOR: &str = env!("CARGO_PKG_AUTHORS");

/// My raytracer, based on the book series on the interwebs.
#[derive(Parser)]
#[clap(version = CRATE_VERSION, author = CRATE_AUTHOR)]
struct Opts {
    #[clap(subcommand)]
    scene: Scene,
    #[clap(long, short, default_value = "400")]
    width: u32,
    #[clap(long, short, default_value = "1.7777778")]
    aspect_ratio: f64,
    #[clap(long, short, default_value = "100")]
    samples_per_pixel: u32,
    /// cpu = the rayon renderer of raytracer_weekend_lib; cuda = librtw_cuda.so (B200)            // +
    #[clap(long, default_value = "cpu")] //                                                          +
    backend: String, //                                                                              +
    /// Philox seed of the cuda backend (the cpu backend keeps thread_rng())                       // +
    #[clap(long, default_value = "0")] //                                                            +
    seed: u64, //                                                                                    +
    /// CUDA device ordinal                                                                        // +
    #[clap(long, default_value = "0")] //                                                            +
    device: i32, //                                                                                  +
}

fn write_png(all_pixels: &[Pixel], image_width: u32, image_height: u32, samples_per_pixel: u32, frame_no: usize) {
    // main.rs:66-94, unchanged
    let mut image = RgbImage::new(image_width, image_height);
    image.pixels_mut().zip(all_pixels.iter()).for_each(|(img_pixel, render_pixel)| {
        let color = render_pixel.color;
        let scale = 1.0 / samples_per_pixel as f32;
        let r = (scale * color.x()).sqrt();
        let g = (scale * color.y()).sqrt();
        let b = (scale * color.z()).sqrt();
        let ir = (255.999 * r.clamp(0.0, 0.999)) as u8;
        let ig = (255.999 * g.clamp(0.0, 0.999)) as u8;
        let ib = (255.999 * b.clamp(0.0, 0.999)) as u8;
        *img_pixel = Rgb([ir, ig, ib]);
    });
    image.save(&format!("render/image_{:04}.png", frame_no)).unwrap();
}

fn main() {
    let opts: Opts = Opts::parse();

    let image_width = opts.width;
    let aspect_ratio = opts.aspect_ratio;
    let image_height = (image_width as f64 / aspect_ratio).round() as u32;
    let samples_per_pixel = opts.samples_per_pixel;
    let pixel_count = (image_width * image_height) as u64;

    let (world, cams, background) = opts.scene.generate((image_width as f32) / (image_height as f32), &mut thread_rng());

    if opts.backend == "cuda" {
        // + flatten once, build the LBVH on the GPU once, render every camera over the resident scene; frame n is
        // + tonemapped and written while frame n+1 renders (rtw_render_frames)
        #[cfg(feature = "cuda")]
        {
            use raytracer_weekend_cuda::{render_params, upload_world};
            let mut scene = upload_world(&world, opts.device).expect("cuda backend: flatten / build failed");
            let params = render_params(background, image_width, image_height, samples_per_pixel, opts.seed);
            scene
                .render_frames(&cams, &params, |frame_no, all_pixels, _stats| {
                    write_png(&all_pixels, image_width, image_height, samples_per_pixel, frame_no as usize);
                    true
                })
                .expect("cuda backend: render failed");
            return;
        }
        #[cfg(not(feature = "cuda"))]
        panic!("console_app was built without the `cuda` feature");
    }

    let overall_progress = ProgressBar::new(cams.len() as u64).with_style(
        ProgressStyle::default_bar().template("[{elapsed_precise} / {eta_precise}] {wide_bar} {pos:>7}/{len:7} ({per_sec}"),
    );
    for (frame_no, cam) in cams.iter().progress_with(overall_progress).enumerate() {
        let raytracer = Raytracer::new(&world, &cam, background, image_width, image_height, samples_per_pixel);
        let frame_progress = ProgressBar::new(pixel_count).with_style(
            ProgressStyle::default_bar().template("[{elapsed_precise} / {eta_precise}] {wide_bar} {pos:>7}/{len:7} ({per_sec})"),
        );
        frame_progress.set_draw_delta(pixel_count / 100);
        let all_pixels: Vec<_> = raytracer.render().progress_with(frame_progress).collect();
        write_png(&all_pixels, image_width, image_height, samples_per_pixel, frame_no);
    }
}
